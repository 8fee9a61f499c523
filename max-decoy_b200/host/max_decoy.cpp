// max_decoy -- native host of the B200 identification hot path: the reference's command line (src/main.rs:213-496) on top
// of the C ABI of include/maxdecoy.h.  The reference's host is compiled Rust (src/main.rs, src/proteomic/tasks); no Rust
// toolchain exists in the build image, so the host side above the C ABI is C++.  Nothing here computes on the data path:
// readers fill the ABI's structure-of-arrays inputs, one md_* call does the work on the GPU, writers format the results.
//
//   max_decoy digest -i db.fasta [-c 2 -l 5 -h 50] [-o dir]              (tasks/digestion.rs:43-136)
//   max_decoy identification -m mods.csv -s run.mzML|run.mgf --fasta db.fasta [-n 0 -d 1000 -l 5 -u 5
//             --fragmentation-tolerance 0.02 -r "<comet revision>"] [-o dir]   (tasks/identification.rs:27-370)
//   max_decoy decoy-generation -m mods.csv -p <mass Da> [-n 0 -d 1000 -l 5 -u 5] [--fasta db.fasta]
//   max_decoy amino-acid-substitution -m mods.csv -s A -d B               (src/main.rs:140-180)
//   max_decoy sequence-mass -s SEQUENCE                                   (tasks/sequence_mass.rs:24-27)
//
// State between `digest` and `identification`: the reference keeps it in PostgreSQL; here `identification` digests the
// FASTA given with --fasta into the in-HBM index (a human proteome takes well under a second) and identifies the whole
// spectrum file in ONE call instead of one process per spectrum.
#include <zlib.h>

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/maxdecoy.h"

namespace {

[[noreturn]] void die(const std::string& msg) {
  std::fprintf(stderr, "max_decoy: %s\n", msg.c_str());
  std::exit(1);
}

std::string read_file(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) die("cannot open " + path);
  std::ostringstream ss;
  ss << in.rdbuf();
  return ss.str();
}
void write_file(const std::string& path, const std::string& text) {
  std::ofstream out(path, std::ios::binary);
  if (!out) die("cannot write " + path);
  out << text;
}
std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) a++;
  while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}
std::vector<std::string> lines_of(const std::string& text) {
  std::vector<std::string> out;
  size_t p = 0;
  while (p <= text.size()) {
    size_t q = text.find('\n', p);
    if (q == std::string::npos) { if (p < text.size()) out.push_back(text.substr(p)); break; }
    out.push_back(text.substr(p, q - p));
    p = q + 1;
  }
  return out;
}

// Rust's `{}` for an f64: shortest round-trip digits, no trailing ".0"
std::string rust_f64(double x) {
  if (x == std::floor(x) && std::fabs(x) < 1e15) { char b[32]; std::snprintf(b, sizeof b, "%lld", (long long)x); return b; }
  char buf[64];
  auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);
  return std::string(buf, r.ptr);
}

// ---------------------------------------------------------------------------------------------- inputs
struct Fasta { std::vector<std::string> headers, sequences; };
// FastaDigester::process_file line loop (utility/input_file_digester/fasta_digester.rs:69-117): lines are trimmed; '>'
// starts a protein; lines before the first header end up in the first protein (reference quirk).
Fasta read_fasta(const std::string& path) {
  Fasta f;
  std::string header, cur;
  for (auto& raw : lines_of(read_file(path))) {
    std::string line = trim(raw);
    if (line.empty() || line[0] != '>') { cur += line; continue; }
    if (!header.empty()) { f.headers.push_back(header); f.sequences.push_back(cur); cur.clear(); }
    header = line;
  }
  f.headers.push_back(header); f.sequences.push_back(cur);
  return f;
}

struct Mod { std::string accession, name; char position; bool is_fix; char aa; int64_t mono; };
// Modification::create_from_csv_file (models/amino_acids/modification.rs:57-76,115-130): 6 columns, first row = header
std::vector<Mod> read_mods(const std::string& path) {
  std::vector<Mod> out;
  auto ls = lines_of(read_file(path));
  for (size_t i = 1; i < ls.size(); i++) {
    if (trim(ls[i]).empty()) continue;
    std::vector<std::string> f;
    std::stringstream ss(ls[i]);
    for (std::string cell; std::getline(ss, cell, ',');) f.push_back(trim(cell));
    if (f.size() != 6) die("modification csv: row has wrong length");   // modification.rs:59-61
    Mod m;
    m.accession = f[0]; for (auto& c : m.accession) c = (char)std::tolower((unsigned char)c);   // modification.rs:48
    m.name = f[1];
    m.position = (char)std::toupper((unsigned char)f[2][0]);
    m.is_fix = std::atoi(f[3].c_str()) > 0;
    m.aa = (char)std::toupper((unsigned char)f[4][0]);
    volatile double v = std::strtod(f[5].c_str(), nullptr) * 1000000.0;   // mass::convert_mass_to_int: truncation
    m.mono = (int64_t)v;
    out.push_back(m);
  }
  return out;
}
std::vector<md_modification> to_abi(const std::vector<Mod>& mods) {
  std::vector<md_modification> out(mods.size());
  for (size_t i = 0; i < mods.size(); i++) {
    std::memset(&out[i], 0, sizeof(out[i]));
    std::strncpy(out[i].accession, mods[i].accession.c_str(), sizeof(out[i].accession) - 1);
    std::strncpy(out[i].name, mods[i].name.c_str(), sizeof(out[i].name) - 1);
    out[i].position = (uint8_t)mods[i].position; out[i].is_fix = mods[i].is_fix; out[i].amino_acid = (uint8_t)mods[i].aa;
    out[i].mono_mass = mods[i].mono;
  }
  return out;
}

struct SpectraSoA {
  std::vector<double> pmz; std::vector<uint8_t> charge; std::vector<uint64_t> off{0}; std::vector<double> mz; std::vector<float> inten;
  std::vector<std::string> spectrum_id, scan_id;
  void push(double m, int z, std::vector<std::pair<double, float>>& peaks, const std::string& sid, const std::string& scan) {
    std::sort(peaks.begin(), peaks.end());
    pmz.push_back(m); charge.push_back((uint8_t)z);
    for (auto& p : peaks) { mz.push_back(p.first); inten.push_back(p.second); }
    off.push_back(mz.size()); spectrum_id.push_back(sid); scan_id.push_back(scan);
  }
  md_spectra abi() const {
    md_spectra s;
    s.n = (uint32_t)pmz.size(); s.precursor_mz = pmz.data(); s.charge = charge.data(); s.spectrum_id = nullptr;
    s.peak_off = off.data(); s.peak_mz = mz.data(); s.peak_intensity = inten.data();
    return s;
  }
};

SpectraSoA read_mgf(const std::string& path) {
  SpectraSoA S;
  std::vector<std::pair<double, float>> peaks; double pm = 0; int z = 0; bool inside = false; int n = 0;
  for (auto& raw : lines_of(read_file(path))) {
    std::string line = trim(raw);
    if (line.empty()) continue;
    if (line == "BEGIN IONS") { peaks.clear(); pm = 0; z = 0; inside = true; }
    else if (line == "END IONS") { n++; S.push(pm, z ? z : 2, peaks, "scan=" + std::to_string(n), std::to_string(n)); inside = false; }
    else if (inside && line.find('=') != std::string::npos && !std::isdigit((unsigned char)line[0])) {
      std::string k = line.substr(0, line.find('=')), v = line.substr(line.find('=') + 1);
      if (k == "PEPMASS") pm = std::strtod(v.c_str(), nullptr);
      else if (k == "CHARGE") z = std::atoi(v.c_str());
    } else if (inside) {
      char* e = nullptr; double a = std::strtod(line.c_str(), &e); double b = std::strtod(e, nullptr);
      peaks.push_back({a, (float)(b == 0.0 && *e == 0 ? 1.0 : b)});
    }
  }
  return S;
}

std::string attr(const std::string& tag, const char* name) {
  std::string key = std::string(name) + "=\"";
  size_t p = tag.find(key);
  if (p == std::string::npos) return "";
  p += key.size();
  size_t q = tag.find('"', p);
  return tag.substr(p, q - p);
}
std::vector<uint8_t> base64_decode(const std::string& s) {
  static int8_t T[256]; static bool init = false;
  if (!init) { std::memset(T, -1, sizeof T); const char* a = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"; for (int i = 0; i < 64; i++) T[(uint8_t)a[i]] = (int8_t)i; init = true; }
  std::vector<uint8_t> out; uint32_t acc = 0; int bits = 0;
  for (unsigned char c : s) { if (T[c] < 0) continue; acc = (acc << 6) | (uint32_t)T[c]; bits += 6; if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); } }
  return out;
}
// MzMlReader::get_ms_two_spectra + Spectrum::new (utility/mz_ml/mz_ml_reader.rs:24-100, spectrum.rs:33-103), plus the peak
// arrays (m/z MS:1000514, intensity MS:1000515; 32/64-bit float; optional zlib) the GPU scorer needs.
SpectraSoA read_mzml(const std::string& path) {
  SpectraSoA S;
  const std::string text = read_file(path);
  size_t p = 0;
  while ((p = text.find("<spectrum ", p)) != std::string::npos) {
    size_t e = text.find("</spectrum>", p);
    if (e == std::string::npos) break;
    const std::string sp = text.substr(p, e - p);
    p = e;
    const std::string open = sp.substr(0, sp.find('>'));
    const std::string sid = attr(open, "id");
    int level = 0; double mz = -1; int z = -1;
    std::vector<double> am, ai;
    size_t q = 0;
    const size_t sel0 = sp.find("<selectedIon>"), sel1 = sp.find("</selectedIon>");
    while ((q = sp.find("<cvParam", q)) != std::string::npos) {
      const std::string tag = sp.substr(q, sp.find('>', q) - q);
      const std::string name = attr(tag, "name");
      if (name == "ms level") level = std::atoi(attr(tag, "value").c_str());
      if (sel0 != std::string::npos && q > sel0 && q < sel1) {
        if (name == "selected ion m/z") mz = std::strtod(attr(tag, "value").c_str(), nullptr);
        if (name == "charge state") z = std::atoi(attr(tag, "value").c_str());
      }
      q += 8;
    }
    if (level != 2) continue;                                   // MzMlReader::is_ms_two_spectrum
    if (mz < 0 || z <= 0 || z > 255 || sid.empty()) die("mzML: MS2 spectrum without selected ion m/z, charge state or id: " + sid);
    q = 0;
    while ((q = sp.find("<binaryDataArray", q)) != std::string::npos) {
      size_t qe = sp.find("</binaryDataArray>", q);
      const std::string bda = sp.substr(q, qe - q);
      q = qe;
      const bool f32 = bda.find("MS:1000521") != std::string::npos || bda.find("32-bit float") != std::string::npos;
      const bool zl = bda.find("MS:1000574") != std::string::npos || bda.find("zlib compression") != std::string::npos;
      const bool is_mz = bda.find("MS:1000514") != std::string::npos, is_int = bda.find("MS:1000515") != std::string::npos;
      size_t b0 = bda.find("<binary>");
      if (b0 == std::string::npos || (!is_mz && !is_int)) continue;
      std::vector<uint8_t> raw = base64_decode(bda.substr(b0 + 8, bda.find("</binary>") - b0 - 8));
      if (zl && !raw.empty()) {
        std::vector<uint8_t> out(raw.size() * 8 + 1024);
        for (;;) {
          uLongf n = out.size();
          int rc = uncompress(out.data(), &n, raw.data(), raw.size());
          if (rc == Z_OK) { out.resize(n); break; }
          if (rc != Z_BUF_ERROR) die("mzML: zlib error in " + sid);
          out.resize(out.size() * 2);
        }
        raw.swap(out);
      }
      std::vector<double>& dst = is_mz ? am : ai;
      if (f32) { dst.resize(raw.size() / 4); for (size_t i = 0; i < dst.size(); i++) { float v; std::memcpy(&v, &raw[4 * i], 4); dst[i] = v; } }
      else { dst.resize(raw.size() / 8); if (!dst.empty()) std::memcpy(dst.data(), raw.data(), dst.size() * 8); }
    }
    std::vector<std::pair<double, float>> peaks;
    for (size_t i = 0; i < std::min(am.size(), ai.size()); i++) peaks.push_back({am[i], (float)ai[i]});
    std::string scan;
    size_t sc = sid.find("scan=");
    if (sc != std::string::npos) { sc += 5; while (sc < sid.size() && std::isdigit((unsigned char)sid[sc])) scan += sid[sc++]; }
    S.push(mz, z, peaks, sid, scan);
  }
  return S;
}
SpectraSoA read_spectra(const std::string& path) {
  const bool mgf = path.size() > 4 && (path.substr(path.size() - 4) == ".mgf" || path.substr(path.size() - 4) == ".MGF");
  return mgf ? read_mgf(path) : read_mzml(path);
}

// ---------------------------------------------------------------------------------------------- outputs
const char* kAlphabet = MD_ALPHABET;
const std::map<char, std::string> kNames = {   // models/amino_acids/amino_acid.rs:7-35
    {'A', "Alanine"}, {'R', "Arginine"}, {'N', "Asparagine"}, {'D', "Aspartic acid"}, {'C', "Cysteine"}, {'E', "Glutamic acid"},
    {'Q', "Glutamine"}, {'G', "Glycine"}, {'H', "Histidine"}, {'J', "Isoleucine or Leucine"}, {'K', "Lysine"}, {'M', "Methionine"},
    {'F', "Phenylalanine"}, {'P', "Proline"}, {'O', "Pyrrolysine"}, {'S', "Serine"}, {'T', "Threonine"}, {'U', "Selenocysteine"},
    {'V', "Valine"}, {'W', "Tryptophan"}, {'Y', "Tyrosine"}};

// ModifiedPeptide::get_modification_summary_for_header (models/peptides/modified_peptide.rs:606-659): the fixed modification
// where its position allows (anywhere; N / C: first / last residue only), the variable one where the mask says
std::string mod_summary(const std::string& seq, const std::vector<Mod>& mods, uint64_t var_mask) {
  std::map<std::string, int> counts;
  for (size_t i = 0; i < seq.size(); i++) {
    const Mod* fix = nullptr; const Mod* var = nullptr;
    for (auto& m : mods) if (m.is_fix && m.aa == seq[i]) fix = &m;
    if (fix && !(fix->position == 'A' || (fix->position == 'N' && i == 0) || (fix->position == 'C' && i + 1 == seq.size()))) fix = nullptr;
    if ((var_mask >> i) & 1) for (auto& m : mods) if (!m.is_fix && m.aa == seq[i]) var = &m;
    if (var && fix && var->position == fix->position) var = nullptr;
    if (fix) counts[fix->accession + "|" + fix->name]++;
    if (var) counts[var->accession + "|" + var->name]++;
  }
  std::string out;
  for (auto& kv : counts) out += "(" + std::to_string(kv.second) + "|" + kv.first + ")";
  return out;
}
std::string lower(std::string s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); return s; }

// comet_parameter::new (utility/comet_parameter.rs:96-124); the fixed part as data (Comet's own parameter names)
std::string comet_params(const std::string& revision, const std::vector<Mod>& mods, const std::string& fasta_path, size_t n_entries, int nvar,
                         double frag_tol, int64_t lppm, int64_t uppm) {
  static const char* const groups[][12][2] = {
      {{"decoy_search", "0"}, {"peff_format", "0"}, {"peff_obo", ""}},
      {{"num_threads", "0"}},
      {{"peptide_mass_units", "2"}, {"mass_type_parent", "1"}, {"mass_type_fragment", "1"}, {"precursor_tolerance_type", "1"}, {"isotope_error", "3"}},
      {{"search_enzyme_number", "1"}, {"num_enzyme_termini", "2"}, {"allowed_missed_cleavage", "2"}},
      {{"max_variable_mods_in_peptide", "5"}, {"require_variable_mod", "0"}},
      {{"theoretical_fragment_ions", "1"}, {"use_A_ions", "0"}, {"use_B_ions", "1"}, {"use_C_ions", "0"}, {"use_X_ions", "0"}, {"use_Y_ions", "1"},
       {"use_Z_ions", "0"}, {"use_NL_ions", "0"}},
      {{"output_sqtstream", "0"}, {"output_sqtfile", "0"}, {"output_txtfile", "1"}, {"output_pepxmlfile", "0"}, {"output_percolatorfile", "0"},
       {"print_expect_score", "1"}, {"show_fragment_ions", "0"}},
      {{"sample_enzyme_number", "1"}},
      {{"scan_range", "0 0"}, {"precursor_charge", "0 0"}, {"override_charge", "0"}, {"ms_level", "2"}, {"activation_method", "ALL"}},
      {{"digest_mass_range", "600.0 5000.0"}, {"skip_researching", "1"}, {"max_fragment_charge", "3"}, {"max_precursor_charge", "6"},
       {"nucleotide_reading_frame", "0"}, {"clip_nterm_methionine", "0"}, {"spectrum_batch_size", "0"}, {"decoy_prefix", "DECOY_"},
       {"equal_I_and_L", "1"}, {"output_suffix", ""}, {"mass_offsets", ""}},
      {{"minimum_peaks", "10"}, {"minimum_intensity", "0"}, {"remove_precursor_peak", "0"}, {"remove_precursor_tolerance", "1.5"},
       {"clear_mz_range", "0.0 0.0"}},
      {{"add_Cterm_peptide", "0.0"}, {"add_Nterm_peptide", "0.0"}, {"add_Cterm_protein", "0.0"}, {"add_Nterm_protein", "0.0"}},
      {{"fragment_bin_offset", "0"}}};
  std::string t = revision + "\n\n# Comet MS/MS search engine parameters file.\n# Everything following the '#' symbol is treated as a comment.\n\n";
  for (auto& g : groups) {
    for (auto& kv : g) { if (!kv[0]) break; t += std::string(kv[0]) + (kv[1][0] ? std::string(" = ") + kv[1] : " =") + "\n"; }
    t += "\n";
  }
  char buf[64]; std::snprintf(buf, sizeof buf, "%.4f", (double)std::max(lppm, uppm));
  t += std::string("peptide_mass_tolerance = ") + buf + "\n";
  t += "fragment_bin_tol = " + rust_f64(frag_tol) + "\n";
  t += "num_results = " + std::to_string(n_entries) + "\nnum_output_lines = " + std::to_string(n_entries) + "\n";
  t += "database_name = " + fasta_path + "\n";
  std::vector<Mod> sorted = mods;
  std::sort(sorted.begin(), sorted.end(), [](const Mod& a, const Mod& b) { return a.aa < b.aa; });
  bool has_j = false;
  for (auto& m : sorted) if (m.is_fix) {
    if (m.aa != 'J') t += std::string("add_") + m.aa + "_" + lower(kNames.count(m.aa) ? kNames.at(m.aa) : "unknown amino acid") + " = " + rust_f64(m.mono / 1000000.0) + "\n";
    else { has_j = true; t += "add_J_user_amino_acid = " + rust_f64((113084060 + m.mono) / 1000000.0) + "\n"; }
  }
  if (!has_j) t += "add_J_user_amino_acid = 113.08406\n";
  int num = 1;
  for (auto& m : sorted) if (!m.is_fix && num <= 9) {
    const int dist = m.position == 'A' ? -1 : 0, term = m.position == 'A' ? 0 : (m.position == 'C' ? 3 : 2);
    t += "variable_mod0" + std::to_string(num++) + " = " + rust_f64(m.mono / 1000000.0) + " " + m.aa + " 0 " + std::to_string(nvar) + " " + std::to_string(dist) +
         " " + std::to_string(term) + " 0\n";
  }
  static const char* const enz[][5] = {{"0.", "No_enzyme", "0", "-", "-"}, {"1.", "Trypsin", "1", "KR", "P"}, {"2.", "Trypsin/P", "1", "KR", "-"},
                                       {"3.", "Lys_C", "1", "K", "P"}, {"4.", "Lys_N", "0", "K", "-"}, {"5.", "Arg_C", "1", "R", "P"},
                                       {"6.", "Asp_N", "0", "D", "-"}, {"7.", "CNBr", "1", "M", "-"}, {"8.", "Glu_C", "1", "DE", "P"},
                                       {"9.", "PepsinA", "1", "FL", "P"}, {"10.", "Chymotrypsin", "1", "FWYL", "P"}};
  t += "\n[COMET_ENZYME_INFO]\n";
  for (auto& e : enz) { std::snprintf(buf, sizeof buf, "%-4s%-23s%-7s%-11s %s\n", e[0], e[1], e[2], e[3], e[4]); t += buf; }
  return t;
}

// ---------------------------------------------------------------------------------------------- command line
struct Args {
  std::map<std::string, std::string> kv;
  std::string get(const std::string& k, const std::string& dflt = "") const { auto it = kv.find(k); return it == kv.end() ? dflt : it->second; }
  bool has(const std::string& k) const { return kv.count(k) != 0; }
  long num(const std::string& k, long dflt) const { return has(k) ? std::atol(kv.at(k).c_str()) : dflt; }
};
// flags: {short, long} -> canonical name
struct Flag { const char* s; const char* l; const char* name; };
Args parse(int argc, char** argv, int first, const std::vector<Flag>& spec) {
  Args a;
  for (int i = first; i < argc; i++) {
    const std::string f = argv[i];
    const char* name = nullptr;
    for (auto& s : spec) if ((s.s[0] && f == std::string("-") + s.s) || f == std::string("--") + s.l) name = s.name;
    if (!name) die("unknown option " + f);
    if (i + 1 >= argc) die("option " + f + " needs a value");
    a.kv[name] = argv[++i];
  }
  return a;
}

void check(md_ctx* ctx, int rc, const char* what) {
  if (rc != MD_OK) die(std::string(what) + ": " + md_last_error(ctx));
}
md_ctx* make_ctx(const Args& a) {
  md_config cfg{(int32_t)a.num("device", 0), 0};
  md_ctx* ctx = nullptr;
  check(nullptr, md_create(&cfg, &ctx), "md_create");
  return ctx;
}
uint64_t digest_into(md_ctx* ctx, const Fasta& f, uint32_t mc, uint32_t min_len, uint32_t max_len) {
  std::vector<uint8_t> res; std::vector<uint64_t> off{0};
  for (auto& s : f.sequences) { res.insert(res.end(), s.begin(), s.end()); off.push_back(res.size()); }
  md_digest_params p{mc, min_len, max_len};
  uint64_t n = 0;
  check(ctx, md_digest(ctx, res.data(), off.data(), (uint32_t)f.sequences.size(), &p, &n), "md_digest");
  return n;
}
std::string count_columns(const int16_t* c) {   // db/schema.sql:20-40: r n d c e q g h j k m f p o s t u v w y, a last
  std::string out;
  for (const char* p = "RNDCEQGHJKMFPOSTUVWYA"; *p; p++) out += "," + std::to_string(c[std::strchr(kAlphabet, *p) - kAlphabet]);
  return out;
}

int cmd_digest(int argc, char** argv) {
  Args a = parse(argc, argv, 2, {{"i", "input-file", "in"}, {"f", "format", "format"}, {"t", "thread-count", "threads"}, {"c", "number-of-missed-cleavages", "mc"},
                                 {"l", "minimum-peptide_length", "min"}, {"h", "maximum-peptide_length", "max"}, {"e", "enzym-name", "enzym"}, {"o", "out", "out"},
                                 {"", "device", "device"}});
  if (!a.has("in")) die("digest: -i/--input-file is required");
  const long mc = a.num("mc", 2), mn = a.num("min", 5), mx = a.num("max", 50);
  if (mc > 60 || mx > 60) die("maximum peptide length and missed cleavages must be <= 60 (tasks/digestion.rs:74,101)");
  Fasta f = read_fasta(a.get("in"));
  md_ctx* ctx = make_ctx(a);
  const uint64_t n = digest_into(ctx, f, (uint32_t)mc, (uint32_t)mn, (uint32_t)mx);
  md_peptide_table t;
  check(ctx, md_peptides_export(ctx, &t), "md_peptides_export");
  const std::string dir = a.get("out", "digest_out");
  mkdir(dir.c_str(), 0777);
  std::string pep, assoc, prot;
  for (uint64_t k = 0; k < t.n; k++) {
    const std::string s((const char*)t.seq + t.seq_off[k], t.seq_off[k + 1] - t.seq_off[k]);
    pep += std::to_string(k + 1) + "," + s + "," + std::to_string(s.size()) + "," + std::to_string(t.missed_cleavages[k]) + "," + std::to_string(t.weight[k]) +
           count_columns(t.counts + k * MD_ALPHABET_SIZE) + "\n";
    for (uint64_t j = t.assoc_off[k]; j < t.assoc_off[k + 1]; j++) assoc += std::to_string(k + 1) + "," + std::to_string(t.assoc_protein[j] + 1) + "\n";
  }
  write_file(dir + "/peptides.csv", pep);
  write_file(dir + "/peptides_proteins.csv", assoc);
  md_peptide_table_free(&t);
  std::printf("%zu proteins, %llu unique peptides -> %s\n", f.sequences.size(), (unsigned long long)n, dir.c_str());
  md_destroy(ctx);
  return 0;
}

int cmd_identification(int argc, char** argv) {
  Args a = parse(argc, argv, 2, {{"m", "modification-file", "mods"}, {"s", "spectrum-file", "spectra"}, {"n", "max-number-of-variable-modification-per-peptide", "nvar"},
                                 {"d", "number-of-decoys", "decoys"}, {"l", "lower-mass-tolerance", "lower"}, {"u", "upper-mass-tolerance", "upper"},
                                 {"", "fragmentation-tolerance", "fragtol"}, {"t", "thread-count", "threads"}, {"", "max-time-for-decoy-generation", "maxtime"},
                                 {"r", "comet-revision", "rev"}, {"", "fasta", "fasta"}, {"c", "number-of-missed-cleavages", "mc"}, {"o", "out", "out"},
                                 {"", "seed", "seed"}, {"", "decoy-mode", "mode"}, {"", "top-k", "topk"}, {"", "device", "device"}, {"", "stored-decoys", "stored"}, {"", "variable-mode", "varmode"},
                                 {"", "rank", "rank"}, {"", "nranks", "nranks"}, {"", "comm-file", "commfile"}});
  if (!a.has("mods") || !a.has("spectra") || !a.has("fasta")) die("identification: -m, -s and --fasta are required");
  const std::vector<Mod> mods = read_mods(a.get("mods"));
  const Fasta f = read_fasta(a.get("fasta"));
  const SpectraSoA S = read_spectra(a.get("spectra"));
  const int nvar = (int)a.num("nvar", 0);
  md_ctx* ctx = make_ctx(a);
  digest_into(ctx, f, (uint32_t)a.num("mc", 2), 5, 50);
  auto am = to_abi(mods);
  check(ctx, md_set_modifications(ctx, am.data(), (uint32_t)am.size(), (uint32_t)nvar), "md_set_modifications");
  {
    const std::string vm = a.get("varmode", "reference");
    if (vm != "reference" && vm != "expanded") die("identification: --variable-mode must be reference or expanded");
    check(ctx, md_set_variable_mode(ctx, vm == "expanded" ? MD_VARMOD_EXPANDED : MD_VARMOD_REFERENCE), "md_set_variable_mode");
  }
  if (a.has("stored")) {
    // the `decoys` table as CSV (id, aa_sequence, ...) or one sequence per line: reused before new decoys are generated
    // (tasks/identification.rs:259-283)
    std::ifstream in(a.get("stored"));
    if (!in) die("identification: cannot read " + a.get("stored"));
    std::string line, blob; std::vector<uint64_t> off{0};
    while (std::getline(in, line)) {
      std::string q = line;
      const size_t c0 = line.find(',');
      if (c0 != std::string::npos) { const size_t c1 = line.find(',', c0 + 1); q = line.substr(c0 + 1, (c1 == std::string::npos ? line.size() : c1) - c0 - 1); }
      std::string t;
      for (char ch : q) if (ch != '"' && ch != ' ' && ch != '\r' && ch != '\t') t.push_back(ch);
      bool ok = !t.empty();
      for (char ch : t) if (ch < 'A' || ch > 'Z') ok = false;
      if (!ok) continue;
      blob += t; off.push_back(blob.size());
    }
    check(ctx, md_decoy_store_set(ctx, (const uint8_t*)blob.data(), off.data(), off.size() - 1), "md_decoy_store_set");
  }
  check(ctx, md_index_build(ctx), "md_index_build");
  md_search_params p;
  std::memset(&p, 0, sizeof p);
  p.lower_ppm = a.num("lower", 5); p.upper_ppm = a.num("upper", 5); p.fragment_tolerance = std::strtod(a.get("fragtol", "0.02").c_str(), nullptr);
  p.n_decoys = (uint32_t)a.num("decoys", 1000); p.decoy_mode = (int32_t)a.num("mode", 0); p.seed = (uint64_t)a.num("seed", 0);
  p.top_k = (uint32_t)a.num("topk", 5); p.min_peaks = 10; p.max_fragment_charge = 3; p.keep_decoys = 1;
  // ---- multi-GPU: one process per GPU (--rank R --nranks N --comm-file PATH, --device = the rank's GPU); the spectra are
  //      sorted by neutral precursor mass and dealt in blocks of 32 over the ranks (every rank sees the same mass mix), the
  //      index is replicated (every rank digested the same FASTA), the PSM tables are gathered with md_gather_psms
  const int rank = (int)a.num("rank", 0), nranks = (int)a.num("nranks", 1);
  if (nranks < 1 || rank < 0 || rank >= nranks) die("identification: --rank / --nranks out of range");
  if (nranks > 1) {
    if (!a.has("commfile")) die("identification: --nranks > 1 needs --comm-file (rank 0 leaves the communicator id there)");
    const std::string cf = a.get("commfile");
    uint8_t id[MD_COMM_ID_BYTES];
    if (rank == 0) {
      check(nullptr, md_comm_unique_id(id), "md_comm_unique_id");
      write_file(cf + ".tmp", std::string((const char*)id, sizeof id));
      if (std::rename((cf + ".tmp").c_str(), cf.c_str()) != 0) die("cannot write " + cf);
    } else {
      std::string got;
      for (int tries = 0; tries < 1200 && got.size() != sizeof id; tries++) {
        std::ifstream in(cf, std::ios::binary);
        if (in) { std::ostringstream ss; ss << in.rdbuf(); got = ss.str(); }
        if (got.size() != sizeof id) usleep(100000);
      }
      if (got.size() != sizeof id) die("identification: no communicator id at " + cf);
      std::memcpy(id, got.data(), sizeof id);
    }
    check(ctx, md_comm_init(ctx, rank, nranks, id), "md_comm_init");
    if (rank == 0) std::remove(cf.c_str());
  }
  const uint32_t n_all = (uint32_t)S.pmz.size();
  constexpr uint32_t kDeal = 32;
  std::vector<uint32_t> mine;    // this rank's spectra (global ordinals), ascending in neutral mass
  {
    std::vector<uint32_t> order(n_all);
    for (uint32_t i = 0; i < n_all; i++) order[i] = i;
    if (nranks > 1)
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        return S.pmz[x] * S.charge[x] - 1.007276 * S.charge[x] < S.pmz[y] * S.charge[y] - 1.007276 * S.charge[y]; });
    for (uint32_t b = 0; b * kDeal < n_all; b++)
      if ((int)(b % (uint32_t)nranks) == rank) for (uint32_t i = b * kDeal; i < std::min(n_all, (b + 1) * kDeal); i++) mine.push_back(order[i]);
  }
  const uint32_t n_blocks = (n_all + kDeal - 1) / kDeal, rows_per_rank = ((n_blocks + nranks - 1) / nranks) * kDeal;
  std::vector<md_psm> my_psms((size_t)rows_per_rank * p.top_k);
  for (auto& r : my_psms) { std::memset(&r, 0, sizeof r); r.spectrum_id = 0xFFFFFFFFu; }   // padding rows
  md_peptide_table pt;
  check(ctx, md_peptides_export(ctx, &pt), "md_peptides_export");
  auto pep_seq = [&](uint64_t id) { return std::string((const char*)pt.seq + pt.seq_off[id - 1], pt.seq_off[id] - pt.seq_off[id - 1]); };
  const std::string dir = a.get("out", "identification_out");
  mkdir(dir.c_str(), 0777);
  const std::string rev = a.get("rev", "# comet_version 2019.01 rev. 4");
  // ---- the spectrum file goes through the library in batches: the decoys of a batch are kept for the per-spectrum FASTA
  //      files, and a batch is sized so that they fit one pass of the workspaces (2^25 decoy slots, 32k spectra); global
  //      spectrum ids keep the decoy RNG streams independent of the batching and of the number of ranks
  const uint32_t batch_max = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(32768, (1ull << 25) / std::max<uint32_t>(1, p.n_decoys)));
  std::map<uint32_t, std::string> csv_rows;     // global spectrum ordinal -> its psms.csv lines
  md_identify_stats total;
  std::memset(&total, 0, sizeof total);
  for (size_t b0 = 0; b0 < mine.size(); b0 += batch_max) {
    const uint32_t bn = (uint32_t)std::min<size_t>(batch_max, mine.size() - b0);
    SpectraSoA B;
    std::vector<uint32_t> gid(bn);
    for (uint32_t k = 0; k < bn; k++) {
      const uint32_t g = mine[b0 + k];
      gid[k] = g;
      B.pmz.push_back(S.pmz[g]); B.charge.push_back(S.charge[g]);
      B.mz.insert(B.mz.end(), S.mz.begin() + S.off[g], S.mz.begin() + S.off[g + 1]);
      B.inten.insert(B.inten.end(), S.inten.begin() + S.off[g], S.inten.begin() + S.off[g + 1]);
      B.off.push_back(B.mz.size());
    }
    md_spectra sp = B.abi();
    sp.spectrum_id = gid.data();
    md_psm* psms = my_psms.data() + b0 * p.top_k;
    md_identify_stats st;
    check(ctx, md_identify(ctx, &sp, &p, psms, &st, nullptr, nullptr), "md_identify");
    total.n_spectra += st.n_spectra; total.n_targets += st.n_targets; total.n_decoys += st.n_decoys; total.n_less_decoys += st.n_less_decoys;
    // the candidate sets that were scored -> the reference's per-spectrum files (tasks/identification.rs:323-368)
    std::vector<md_precursor> pre(bn);
    for (uint32_t i = 0; i < bn; i++) {
      md_precursor_window(B.pmz[i], B.charge[i], p.lower_ppm, p.upper_ppm, &pre[i].mass, &pre[i].lo, &pre[i].hi);
      pre[i].charge = B.charge[i]; pre[i].spectrum_id = gid[i];
    }
    md_candidate_table cand; md_decoy_table dec;
    check(ctx, md_last_decoys_export(ctx, &dec), "md_last_decoys_export");
    check(ctx, md_candidates(ctx, pre.data(), bn, &cand), "md_candidates");
    auto dec_seq = [&](uint64_t i) { return std::string((const char*)dec.seq + dec.seq_off[i], dec.seq_off[i + 1] - dec.seq_off[i]); };
    for (uint32_t s = 0; s < bn; s++) {
      const uint32_t g = gid[s];
      const std::string name = S.scan_id[g].empty() ? S.spectrum_id[g] : S.scan_id[g];
      std::string fasta; std::set<std::string> seen_t, seen_d;
      for (uint64_t i = cand.off[s]; i < cand.off[s + 1]; i++) {
        const std::string q = pep_seq(cand.peptide_id[i]);
        if (!seen_t.insert(q).second) continue;
        const std::string sum = mod_summary(q, mods, cand.var_mask[i]);
        fasta += ">PEPTIDE_" + q + " MaxDecoyId=" + std::to_string(cand.peptide_id[i]) + (sum.empty() ? "" : " ModRes=" + sum) + "\n" + q + "\n";
      }
      for (uint64_t i = dec.off[s]; i < dec.off[s + 1]; i++) {
        const std::string q = dec_seq(i);
        if (!seen_d.insert(q).second) continue;
        const std::string sum = mod_summary(q, mods, dec.var_mask[i]);
        fasta += ">DECOY_" + q + (sum.empty() ? "" : " ModRes=" + sum) + "\n" + q + "\n";
      }
      const std::string fpath = dir + "/" + name + ".fasta";
      write_file(fpath, fasta);
      if (seen_d.size() < p.n_decoys) write_file(dir + "/" + name + ".less_decoys", std::to_string(seen_d.size()));
      write_file(dir + "/" + name + ".comet.params", comet_params(rev, mods, fpath, seen_t.size() + seen_d.size(), nvar, p.fragment_tolerance, p.lower_ppm, p.upper_ppm));
      std::string& csv = csv_rows[g];
      for (uint32_t r = 0; r < p.top_k; r++) {
        const md_psm& row = psms[(size_t)s * p.top_k + r];
        if (!row.rank) continue;
        const std::string q = row.is_decoy ? dec_seq(dec.off[s] + row.candidate) : pep_seq(row.candidate);
        char sc[48]; std::snprintf(sc, sizeof sc, "%.9g", (double)row.score);
        csv += S.spectrum_id[g] + "," + S.scan_id[g] + "," + std::to_string(row.rank) + "," + (row.is_decoy ? "t," : "f,") +
               (row.is_decoy ? "" : std::to_string(row.candidate)) + "," + q + "," + mod_summary(q, mods, row.var_mask) + "," + std::to_string(row.mod_weight) + "," +
               std::to_string(pre[s].mass) + "," + std::to_string(row.charge) + "," + sc + "," + std::to_string(row.n_targets + row.n_decoys) + "\n";
      }
    }
    md_candidate_table_free(&cand); md_decoy_table_free(&dec);
  }
  // ---- PSM rows of this rank (spectrum order); with several ranks every rank leaves its part and rank 0, once the gather
  //      of the PSM tables has brought everybody's rows, checks the parts against the gathered table and joins them
  auto part_path = [&](int r) { return dir + "/psms.rank" + std::to_string(r) + ".csv"; };
  std::string csv;
  for (auto& kv : csv_rows) csv += kv.second;
  if (nranks == 1) write_file(dir + "/psms.csv", csv);
  else {
    write_file(part_path(rank) + ".tmp", csv);
    std::rename((part_path(rank) + ".tmp").c_str(), part_path(rank).c_str());
    std::vector<md_psm> all((size_t)rows_per_rank * p.top_k * nranks);
    check(ctx, md_gather_psms(ctx, my_psms.data(), (uint64_t)rows_per_rank * p.top_k, all.data()), "md_gather_psms");
    if (rank == 0) {
      uint64_t rows_ranked = 0; std::set<uint32_t> seen;
      for (auto& r : all) if (r.spectrum_id != 0xFFFFFFFFu) { seen.insert(r.spectrum_id); if (r.rank) rows_ranked++; }
      if (seen.size() != n_all) die("md_gather_psms: gathered " + std::to_string(seen.size()) + " spectra, expected " + std::to_string(n_all));
      std::map<std::string, std::vector<std::string>> by_spectrum;   // spectrum id -> lines, parts in any order
      uint64_t lines = 0;
      for (int r = 0; r < nranks; r++) {
        std::string text;
        for (int tries = 0; tries < 600; tries++) { std::ifstream in(part_path(r)); if (in) { std::ostringstream ss; ss << in.rdbuf(); text = ss.str(); break; } usleep(100000); }
        for (auto& ln : lines_of(text)) { if (ln.empty()) continue; by_spectrum[ln.substr(0, ln.find(','))].push_back(ln); lines++; }
        std::remove(part_path(r).c_str());
      }
      if (lines != rows_ranked) die("psms.csv: " + std::to_string(lines) + " lines in the ranks' parts, " + std::to_string(rows_ranked) + " ranked rows gathered");
      std::string joined;
      for (uint32_t g = 0; g < n_all; g++) { auto it = by_spectrum.find(S.spectrum_id[g]); if (it != by_spectrum.end()) { for (auto& ln : it->second) joined += ln + "\n"; by_spectrum.erase(it); } }
      write_file(dir + "/psms.csv", joined);
    }
    check(ctx, md_comm_destroy(ctx), "md_comm_destroy");
  }
  std::printf("%s%llu spectra, %llu targets, %llu decoys scored, %llu spectra with fewer decoys than requested -> %s\n",
              nranks > 1 ? ("rank " + std::to_string(rank) + "/" + std::to_string(nranks) + ": ").c_str() : "", (unsigned long long)total.n_spectra,
              (unsigned long long)total.n_targets, (unsigned long long)total.n_decoys, (unsigned long long)total.n_less_decoys, dir.c_str());
  md_peptide_table_free(&pt);
  md_destroy(ctx);
  return 0;
}

int cmd_decoy_generation(int argc, char** argv) {
  Args a = parse(argc, argv, 2, {{"m", "modification-file", "mods"}, {"n", "max-modification-per-decoy", "nvar"}, {"p", "precursor-mass", "mass"},
                                 {"d", "number-of-decoys", "decoys"}, {"l", "lower-mass-tolerance", "lower"}, {"u", "upper-mass-tolerance", "upper"},
                                 {"t", "thread-count", "threads"}, {"", "max-time-for-decoy-generation", "maxtime"}, {"", "fasta", "fasta"}, {"", "seed", "seed"},
                                 {"", "decoy-mode", "mode"}, {"", "device", "device"}});
  if (!a.has("mods") || !a.has("mass")) die("decoy-generation: -m and -p are required");
  const std::vector<Mod> mods = read_mods(a.get("mods"));
  md_ctx* ctx = make_ctx(a);
  Fasta f; if (a.has("fasta")) f = read_fasta(a.get("fasta"));
  digest_into(ctx, f, 2, 5, 50);
  auto am = to_abi(mods);
  check(ctx, md_set_modifications(ctx, am.data(), (uint32_t)am.size(), (uint32_t)a.num("nvar", 0)), "md_set_modifications");
  check(ctx, md_index_build(ctx), "md_index_build");
  volatile double v = std::strtod(a.get("mass").c_str(), nullptr) * 1000000.0;
  md_precursor pr; pr.mass = (int64_t)v; pr.charge = 2; pr.spectrum_id = 0;
  // (the reference passes the ppm integers as absolute limits in swapped order, src/main.rs:126-133; fixed here)
  pr.lo = pr.mass - pr.mass * a.num("lower", 5) / 1000000; pr.hi = pr.mass + pr.mass * a.num("upper", 5) / 1000000;
  md_decoy_table d;
  check(ctx, md_generate_decoys(ctx, &pr, 1, (uint32_t)a.num("decoys", 1000), (int)a.num("mode", 0), (uint64_t)a.num("seed", 0), &d), "md_generate_decoys");
  for (uint64_t i = 0; i < d.n; i++) std::printf("%.*s\n", (int)(d.seq_off[i + 1] - d.seq_off[i]), (const char*)d.seq + d.seq_off[i]);
  md_decoy_table_free(&d);
  md_destroy(ctx);
  return 0;
}

int cmd_substitution(int argc, char** argv) {
  Args a = parse(argc, argv, 2, {{"m", "modification_file", "mods"}, {"s", "source-amino-acid", "src"}, {"d", "destination-amino-acid", "dst"}, {"", "device", "device"}});
  if (!a.has("mods") || !a.has("src") || !a.has("dst")) die("amino-acid-substitution: -m, -s and -d are required");
  const char* ps = std::strchr(kAlphabet, std::toupper((unsigned char)a.get("src")[0]));
  const char* pd = std::strchr(kAlphabet, std::toupper((unsigned char)a.get("dst")[0]));
  if (!ps || !pd) die("amino acids must be among " MD_ALPHABET);
  md_ctx* ctx = make_ctx(a);
  auto am = to_abi(read_mods(a.get("mods")));
  check(ctx, md_set_modifications(ctx, am.data(), (uint32_t)am.size(), 0), "md_set_modifications");
  int64_t map[MD_ALPHABET_SIZE * MD_ALPHABET_SIZE];
  check(ctx, md_substitution_map(ctx, map), "md_substitution_map");
  std::printf("%s\n", rust_f64(map[(ps - kAlphabet) * MD_ALPHABET_SIZE + (pd - kAlphabet)] / 1000000.0).c_str());
  md_destroy(ctx);
  return 0;
}

int cmd_sequence_mass(int argc, char** argv) {
  Args a = parse(argc, argv, 2, {{"s", "sequence", "seq"}});
  if (!a.has("seq")) die("sequence-mass: -s is required");
  const std::string s = a.get("seq");
  std::printf("%s\n", rust_f64(md_sequence_weight((const uint8_t*)s.data(), (uint32_t)s.size()) / 1000000.0).c_str());
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) die("usage: max_decoy {digest|identification|decoy-generation|amino-acid-substitution|sequence-mass} [options]");
  const std::string cmd = argv[1];
  if (cmd == "digest") return cmd_digest(argc, argv);
  if (cmd == "identification") return cmd_identification(argc, argv);
  if (cmd == "decoy-generation") return cmd_decoy_generation(argc, argv);
  if (cmd == "amino-acid-substitution") return cmd_substitution(argc, argv);
  if (cmd == "sequence-mass") return cmd_sequence_mass(argc, argv);
  die("unknown subcommand " + cmd);
}
