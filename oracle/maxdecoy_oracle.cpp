// maxdecoy_oracle.cpp -- CPU restatement of the MaxDecoy identification hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA library: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The product (max-decoy_b200/) never links, imports or executes anything under oracle/.
//
// Parity status
//   * digest, integer masses, precursor windows: PINNED by the reference's own golden vector
//     (P77377 -> 71 peptides, src/proteomic/models/enzyms/tests/digest_enzym.rs:13-92) and by
//     the constants table (models/amino_acids/amino_acid.rs:7-35); see tests/test_oracle_golden.py.
//   * candidate lookup / ModifiedPeptide filter: restated literally (two independent
//     formulations are cross-checked in tests: SQL fan-out enumeration vs the W* window).
//   * decoys: the reference RNG is unseeded (utility/decoy_generator.rs:130) -> property parity
//     only; this file defines the seeded counter-RNG variant the CUDA kernel must match bit-exactly.
//   * scoring: PARITY UNPINNED.  The reference contains no scorer (it writes comet.params and
//     shells out to Comet: tasks/identification.rs:358-368, run_splitup_and_identification.sh:47-60).
//     The score below is a Comet-style fast-xcorr consistent with utility/comet_parameter.rs:6-79,
//     defined in exact integer arithmetic; it is NOT validated against a Comet binary.
//
// All file:line citations are relative to /root/reference/src/proteomic/.
// Build: see oracle/Makefile (-O2 -ffp-contract=off: the window math must not be fused).

#include "../include/maxdecoy.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------
// masses: models/mass/mod.rs:3-8, models/amino_acids/amino_acid.rs:7-35
// ---------------------------------------------------------------------------------------
inline int64_t convert_mass_to_int(double m) { return (int64_t)(m * 1000000.0); }  // mass/mod.rs:6-8

double mono_mass_f64(uint8_t c) {  // amino_acid.rs:7-35 (column 5), get(): :87-117
  switch (c) {
    case 'A': return 71.03711;  case 'B': return 114.53495; case 'R': return 156.10111;
    case 'N': return 114.04293; case 'D': return 115.02694; case 'C': return 103.00919;
    case 'E': return 129.04259; case 'Q': return 128.05858; case 'G': return 57.02146;
    case 'H': return 137.05891; case 'I': return 113.08406; case 'L': return 113.08406;
    case 'J': return 113.08406; case 'K': return 128.09496; case 'M': return 131.04049;
    case 'F': return 147.06841; case 'P': return 97.05276;  case 'O': return 109.0528;
    case 'S': return 87.03203;  case 'T': return 101.04768; case 'U': return 150.95363;
    case 'V': return 99.06841;  case 'W': return 186.07931; case 'X': return 0.0;
    case 'Y': return 163.06333; case 'Z': return 128.55059;
    default: return 0.0;  // unknown -> X (amino_acid.rs:115)
  }
}
inline int64_t residue_mass(uint8_t c) { return convert_mass_to_int(mono_mass_f64(c)); }

const char kAlphabet[] = MD_ALPHABET;  // amino_acid.rs:4-5
int alpha_index(uint8_t c) {
  for (int i = 0; i < MD_ALPHABET_SIZE; i++)
    if ((uint8_t)kAlphabet[i] == c) return i;
  return -1;
}

int64_t sequence_weight(const uint8_t* s, uint32_t len) {  // amino_acid.rs:130-136
  int64_t w = convert_mass_to_int(18.010565);              // neutral_loss.rs:3
  for (uint32_t i = 0; i < len; i++) w += residue_mass(s[i]);
  return w;
}

// identification.rs:203-211; utility/mod.rs:9-11; mass/mod.rs:14-16
void precursor_window(double mz, uint32_t z, int64_t lppm, int64_t uppm, int64_t* P, int64_t* lo,
                      int64_t* hi) {
  const double H = 1.007276;
  double zc = (double)(uint8_t)z;
  double tl = mz / 1000000.0 * (double)lppm;
  double tu = mz / 1000000.0 * (double)uppm;
  volatile double a = mz * zc;  // volatile: keep every product rounded separately
  volatile double b = H * zc;
  *P = convert_mass_to_int(a - b);
  volatile double al = (mz - tl) * zc;
  *lo = convert_mass_to_int(al - b);
  volatile double au = (mz + tu) * zc;
  *hi = convert_mass_to_int(au - b);
}

// hash of a generalized sequence: part of the canonical peptide order (builder-defined).
uint64_t hash64(const uint8_t* s, uint32_t len) {
  uint64_t h = 0xcbf29ce484222325ULL;
  for (uint32_t i = 0; i < len; i++) { h ^= s[i]; h *= 0x100000001b3ULL; }
  h ^= len;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
  return h;
}

// Philox4x32-10 counter RNG (builder-defined; the reference uses an unseeded thread_rng).
struct Philox {
  uint32_t key[2]; uint32_t ctr[4]; uint32_t out[4]; int have;
  Philox(uint64_t seed, uint32_t spectrum_id, uint32_t attempt, uint32_t tag) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0; ctr[1] = attempt; ctr[2] = spectrum_id; ctr[3] = tag; have = 0;
  }
  static void round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  void refill() {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}; uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) { round(c, k); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    for (int i = 0; i < 4; i++) out[i] = c[i];
    ctr[0]++; have = 4;
  }
  uint32_t next() { if (!have) refill(); return out[4 - have--]; }
  uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};
// diagnostics (oracle only, not part of the ABI): attempts, successes, greedy passes, position evaluations
std::atomic<uint64_t> g_stat_attempts(0), g_stat_success(0), g_stat_tries(0), g_stat_evals(0);

const uint32_t kTagRandom = 0x4D444543u;   // "MDEC"
const uint32_t kTagPermute = 0x4D445045u;  // "MDPE"

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
struct ModSet {
  bool set = false;
  uint32_t nvar = 0;                       // -n
  bool has_fix[MD_ALPHABET_SIZE] = {};
  bool has_var[MD_ALPHABET_SIZE] = {};
  int64_t fix[MD_ALPHABET_SIZE] = {};      // fixed delta per letter
  int64_t var[MD_ALPHABET_SIZE] = {};      // variable delta per letter
  // position of the letter's fixed / variable modification: 'A' anywhere, 'N' / 'C' terminus (modification.rs:24-33).
  // A terminal modification sits on the first / last residue only, and only when that residue is its letter
  // (add_modification_at, modified_peptide.rs:421-447; set_variable_modification_at, :339-367).  One fixed and one
  // variable modification per letter (HashMap<char, Modification>, identification.rs:163-171).
  uint8_t fix_pos[MD_ALPHABET_SIZE] = {};
  uint8_t var_pos[MD_ALPHABET_SIZE] = {};
  bool has_terminal = false;
  static bool at(uint8_t pos, uint32_t i, uint32_t L) { return pos == 'A' || (pos == 'N' && i == 0) || (pos == 'C' && i + 1 == L); }
  // does residue i of a sequence of L residues, letter a, carry the letter's fixed modification?
  bool fix_applies(int a, uint32_t i, uint32_t L) const { return has_fix[a] && at(fix_pos[a], i, L); }
  // can it take the letter's variable modification?  (its slot -- side chain, N-terminus, C-terminus -- must exist at this
  // position and must not hold the fixed modification: AlreadyFixModificationInPlace, :311,327,350)
  bool can_var(int a, uint32_t i, uint32_t L) const { return has_var[a] && at(var_pos[a], i, L) && !(has_fix[a] && fix_pos[a] == var_pos[a]); }
  std::vector<int> letters;                // sorted modifiable letters (identification.rs:173-178), alphabet idx
  std::vector<uint8_t> letter_chars;
  int64_t merged(int a) const { return has_var[a] ? var[a] : fix[a]; }  // identification.rs:190-196
  int64_t mprime(int a) const { return residue_mass((uint8_t)kAlphabet[a]) + (has_fix[a] ? fix[a] : 0); }
};

struct Peptides {
  bool ready = false;
  std::vector<std::string> seq;
  std::vector<int64_t> weight;
  std::vector<uint8_t> mc;
  std::vector<int16_t> counts;  // n*21
  std::vector<std::vector<uint32_t>> assoc;
  std::unordered_map<std::string, uint32_t> by_seq;
};

struct Index {
  bool ready = false;
  std::vector<int64_t> key;   // W*
  std::vector<uint32_t> pep;  // 0-based peptide ordinal
};

// Decoys persisted by earlier runs (the `decoys` table, db/schema.sql; Decoy::find_where, models/peptides/decoy.rs:118-153):
// unique sequences in canonical order (weight, hash, bytes) + their index by W* (same key as the peptides).
struct DecoyStore {
  std::vector<std::string> seq;
  std::vector<int64_t> weight;
  std::vector<int16_t> counts;  // n*21
  bool indexed = false;
  std::vector<int64_t> key;
  std::vector<uint32_t> ord;
};

struct Candidate { uint32_t pep; uint64_t mask; int64_t w; };
struct Decoy { std::string seq; uint64_t mask; int64_t weight; int64_t w; uint32_t attempt; };

}  // namespace

struct md_ctx {
  std::string err;
  uint32_t n_threads = 1;
  ModSet mods;
  Peptides peps;
  Index index;
  DecoyStore store;
  int var_mode = 0;                                   // md_varmod_mode
  std::vector<std::pair<int64_t, uint32_t>> fixed;    // MD_VARMOD_EXPANDED: (weight incl. fixed mods, index entry), ascending
  std::vector<std::vector<Decoy>> last_decoys;
  bool have_last_decoys = false;
};

namespace {

thread_local std::string g_err;
int fail(md_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_err = msg;
  return code;
}

// ---------------------------------------------------------------------------------------
// digest: models/enzyms/digest_enzym.rs:61-86, trypsin.rs:29, peptides/peptide.rs:27-37
// ---------------------------------------------------------------------------------------
void split_trypsin(const uint8_t* s, size_t n, std::vector<std::pair<size_t, size_t>>* pieces) {
  // onig split on (?<=[KR])(?!P): cut before position p (0<p<n) when s[p-1] in {K,R} and s[p] != P.
  // A protein ending in K/R yields no trailing empty piece worth keeping (it would only be
  // concatenated as an empty string onto sequences that are emitted anyway).
  size_t start = 0;
  for (size_t p = 1; p < n; p++) {
    if ((s[p - 1] == 'K' || s[p - 1] == 'R') && s[p] != 'P') { pieces->push_back({start, p}); start = p; }
  }
  if (n > 0) pieces->push_back({start, n});
}

struct Occ { std::string g; uint8_t mc; uint32_t prot; };

// ---------------------------------------------------------------------------------------
// ModifiedPeptide filter: models/peptides/modified_peptide.rs:118-159 (from_string, fixed mods),
// :512-543 (try_variable_modifications), utility/combinations/n_choose_k.rs:12-49 (order).
// Terminal (N/C) modifications follow the slot model the reference's add_modification_at / set_variable_modification_at
// spell out (ModSet::fix_applies / can_var); its `modifications` vector bookkeeping for them is not restated (push_modification
// leaves the vector one entry short per terminal modification, :182-208, which makes later index-based calls hit the wrong
// residue or panic -- there is no behaviour to pin there).
// ---------------------------------------------------------------------------------------
struct ModState {  // the mutable working peptide
  std::vector<int> aa;          // alphabet idx, or -1 for letters outside the alphabet
  std::vector<uint8_t> raw;     // letter
  std::vector<uint8_t> mod;     // bit 0: the letter's fixed modification is applied, bit 1: its variable one   (ModifiedPeptide.modifications + n/c_terminus_modification)
  int64_t w;
};

inline bool in_window(int64_t w, int64_t lo, int64_t hi) { return lo <= w && w <= hi; }  // :157-159

void from_string(const ModSet& M, const uint8_t* s, uint32_t len, ModState* st) {
  st->aa.resize(len); st->raw.assign(s, s + len); st->mod.assign(len, 0);
  st->w = convert_mass_to_int(18.010565);
  for (uint32_t i = 0; i < len; i++) {
    int a = alpha_index(s[i]);
    st->aa[i] = a;
    st->w += residue_mass(s[i]);
    if (a >= 0 && M.fix_applies(a, i, len)) { st->w += M.fix[a]; st->mod[i] = 1; }
  }
}

void remove_all_var(const ModSet& M, ModState* st) {  // :369-401
  for (size_t i = 0; i < st->mod.size(); i++)
    if (st->mod[i] & 2) { st->w -= M.var[st->aa[i]]; st->mod[i] &= (uint8_t)~2; }
}

uint64_t var_mask_of(const ModState& st) {
  uint64_t m = 0;
  for (size_t i = 0; i < st.mod.size(); i++) if (st.mod[i] & 2) m |= 1ULL << i;
  return m;
}

// previous k-subset mask in descending numeric order, or 0 if `m` was the smallest (2^k-1).
inline uint64_t prev_combination(uint64_t m) {
  // trailing ones t, then first zero run; standard "previous bit permutation"
  uint64_t t = m & (~m + 1);              // lowest set bit
  if (t != 1) {                           // lowest bit is not bit0: move it down by one
    return (m & ~t) | (t >> 1);
  }
  // m ends in a block of ones at bit 0: ones = trailing ones count
  uint64_t ones = m & ~(m + 1);           // trailing ones block (bits 0..j-1)
  uint64_t rest = m & ~ones;
  if (rest == 0) return 0;                // m == 2^k - 1: smallest
  uint64_t low = rest & (~rest + 1);      // lowest set bit above the block
  int j = __builtin_popcountll(ones);
  // move `low` down by one and pack the j trailing ones right below it
  uint64_t moved = low >> 1;
  uint64_t packed = ((1ULL << j) - 1) * (moved >> j);  // j ones ending just below `moved`
  // (moved >> j) is a power of two 2^(pos-j); times (2^j-1) gives ones at bits pos-j..pos-1
  return (rest & ~low) | moved | packed;
}

bool try_variable(const ModSet& M, ModState* st, int64_t lo, int64_t hi) {  // :512-543
  std::vector<uint32_t> pos;
  for (uint32_t i = 0; i < st->aa.size(); i++)
    if (st->aa[i] >= 0 && M.has_var[st->aa[i]]) pos.push_back(i);
  const uint32_t d = (uint32_t)pos.size();
  for (uint32_t n = 1; n <= M.nvar; n++) {
    if (n > d) continue;
    // NChooseK: masks over d bits, MSB <-> first position, descending from 2^d - 2^(d-n)
    uint64_t mask = (d == 64 ? ~0ULL : ((1ULL << d) - 1)) ^ ((1ULL << (d - n)) - 1);
    while (mask) {
      remove_all_var(M, st);
      for (uint32_t b = 0; b < d; b++) {
        if (!((mask >> (d - 1 - b)) & 1)) continue;
        uint32_t i = pos[b];
        // AlreadyFixModificationInPlace -> continue 'positions (:355,532); a terminal modification away from its terminus
        // falls through set_variable_modification_at without effect (:345-366)
        if (!M.can_var(st->aa[i], i, (uint32_t)st->aa.size())) continue;
        st->mod[i] |= 2; st->w += M.var[st->aa[i]];
      }
      if (in_window(st->w, lo, hi)) return true;
      mask = prev_combination(mask);
    }
  }
  return false;
}

// identification.rs:214-222,374-403: is the count vector of this peptide among the enumerated
// SQL queries?  (k_a < K_a for every modifiable letter, and the incoming lower limit of every
// recursion level is > 0.)
bool fanout_admits(const ModSet& M, const int16_t* counts, int64_t P, int64_t lo) {
  if (M.letters.empty()) return false;  // no modifiable letters -> no queries at all (:375-379)
  int64_t cur_lo = lo;
  for (size_t i = 0; i < M.letters.size(); i++) {
    int a = M.letters[i];
    int16_t K = (int16_t)(P / (residue_mass((uint8_t)kAlphabet[a]) + M.merged(a)));
    if (!(counts[a] < K)) return false;
    if (!(cur_lo > 0)) return false;
    cur_lo -= (int64_t)counts[a] * M.merged(a);
  }
  return true;
}

// MD_VARMOD_EXPANDED (builder-defined): every placement of up to nvar variable modifications whose
// weight lies in the window is a candidate.  Count vectors over the variable letters (alphabetical, last letter
// fastest), per vector the window on the fixed-modification weight, per peptide the placements (last letter fastest,
// each letter's subsets in NChooseK order, n_choose_k.rs:12-49).
void candidates_expanded(const md_ctx* ctx, const md_precursor& pr, std::vector<Candidate>* out) {
  const ModSet& M = ctx->mods; const Index& X = ctx->index; const Peptides& Pp = ctx->peps;
  std::vector<int> letters;                                // variable-modifiable letters by character
  {
    std::vector<uint8_t> chars;
    for (int a = 0; a < MD_ALPHABET_SIZE; a++) if (M.has_var[a] && !(M.has_fix[a] && M.fix_pos[a] == M.var_pos[a])) chars.push_back((uint8_t)kAlphabet[a]);   // (unless its slot holds the fixed modification)
    std::sort(chars.begin(), chars.end());
    for (uint8_t c : chars) letters.push_back(alpha_index(c));
  }
  const size_t nl = letters.size();
  std::vector<uint32_t> k(nl, 0);
  for (;;) {
    uint32_t sum = 0; int64_t shift = 0;
    for (size_t a = 0; a < nl; a++) { sum += k[a]; shift += (int64_t)k[a] * M.var[letters[a]]; }
    if (sum <= M.nvar) {
      const int64_t lo = pr.lo - shift, hi = pr.hi - shift;
      auto b = std::lower_bound(ctx->fixed.begin(), ctx->fixed.end(), std::make_pair(lo, (uint32_t)0));
      for (auto it = b; it != ctx->fixed.end() && it->first <= hi; ++it) {
        const uint32_t p = X.pep[it->second];
        const std::string& q = Pp.seq[p];
        std::vector<std::vector<uint32_t>> pos(nl);
        bool ok = true;
        for (size_t a = 0; a < nl; a++) {
          for (uint32_t i = 0; i < q.size(); i++) if (q[i] == kAlphabet[letters[a]] && ModSet::at(M.var_pos[letters[a]], i, (uint32_t)q.size())) pos[a].push_back(i);   // where the modification's slot exists
          if (pos[a].size() < k[a]) ok = false;
        }
        if (!ok) continue;
        // odometer over the letters' k-subsets: compressed masks, MSB <-> first position, descending
        std::vector<uint64_t> cm(nl), first(nl);
        for (size_t a = 0; a < nl; a++) {
          const uint32_t d = (uint32_t)pos[a].size();
          first[a] = k[a] == 0 ? 0 : (((1ULL << d) - 1) ^ ((1ULL << (d - k[a])) - 1));
          cm[a] = first[a];
        }
        for (;;) {
          uint64_t mask = 0;
          for (size_t a = 0; a < nl; a++) {
            const uint32_t d = (uint32_t)pos[a].size();
            for (uint32_t bix = 0; bix < d; bix++) if ((cm[a] >> (d - 1 - bix)) & 1) mask |= 1ULL << pos[a][bix];
          }
          out->push_back({p, mask, it->first + shift});
          int a = (int)nl - 1;
          for (; a >= 0; a--) {
            if (k[a] == 0) continue;
            const uint64_t nx = prev_combination(cm[a]);
            if (nx) { cm[a] = nx; break; }
            cm[a] = first[a];
          }
          if (a < 0) break;
        }
      }
    }
    int a = (int)nl - 1;
    while (a >= 0 && k[a] >= M.nvar) { k[a] = 0; a--; }
    if (a < 0) break;
    k[a]++;
  }
}

void candidates_for(const md_ctx* ctx, const md_precursor& pr, std::vector<Candidate>* out) {
  if (ctx->var_mode == MD_VARMOD_EXPANDED) { candidates_expanded(ctx, pr, out); return; }
  const ModSet& M = ctx->mods; const Index& X = ctx->index; const Peptides& Pp = ctx->peps;
  size_t b = std::lower_bound(X.key.begin(), X.key.end(), pr.lo) - X.key.begin();
  size_t e = std::upper_bound(X.key.begin(), X.key.end(), pr.hi) - X.key.begin();
  ModState st;
  for (size_t i = b; i < e; i++) {
    uint32_t p = X.pep[i];
    if (!fanout_admits(M, &Pp.counts[(size_t)p * MD_ALPHABET_SIZE], pr.mass, pr.lo)) continue;
    const std::string& s = Pp.seq[p];
    from_string(M, (const uint8_t*)s.data(), (uint32_t)s.size(), &st);  // identification.rs:245
    bool ok = in_window(st.w, pr.lo, pr.hi);
    if (!ok) ok = try_variable(M, &st, pr.lo, pr.hi);                   // :247-249
    if (ok) out->push_back({p, var_mask_of(st), st.w});
  }
}

// ---------------------------------------------------------------------------------------
// decoys
// ---------------------------------------------------------------------------------------
// remove_modification_at (:403-419): whatever residue i carries (side chain and, at the ends, terminus)
inline void remove_mod_at(const ModSet& M, ModState* st, uint32_t i) {
  if (st->mod[i] & 1) st->w -= M.fix[st->aa[i]];
  if (st->mod[i] & 2) st->w -= M.var[st->aa[i]];
  st->mod[i] = 0;
}
inline void replace_at(const ModSet& M, ModState* st, uint32_t i, int c) {  // :470-482 / :493-505
  remove_mod_at(M, st, i);
  st->w -= residue_mass(st->raw[i]);
  st->w += residue_mass((uint8_t)kAlphabet[c]);
  st->aa[i] = c; st->raw[i] = (uint8_t)kAlphabet[c];
  if (M.fix_applies(c, i, (uint32_t)st->aa.size())) { st->w += M.fix[c]; st->mod[i] = 1; }   // add_modification_at (:421-447)
}

// One attempt of DecoyGenerator::generate_decoys' worker loop (utility/decoy_generator.rs:139-187)
// + swap_amino_acids_to_hit_mass_tolerance (modified_peptide.rs:451-508).
// Deviations that make it reproducible (documented in DESIGN.md): counter RNG; ties in the greedy
// step resolved by alphabet order (the reference follows HashMap iteration order); attempts whose
// grown sequence exceeds 60 residues are dropped (the reference fails on VARCHAR(60) at insert).
bool random_attempt(const ModSet& M, const md_precursor& pr, uint64_t seed, uint32_t attempt,
                    const int64_t* delta /*21x21*/, ModState* st) {
  Philox rng(seed, pr.spectrum_id, attempt, kTagRandom);
  st->aa.clear(); st->raw.clear(); st->mod.clear();
  st->w = convert_mass_to_int(18.010565);
  for (;;) {                                            // 'amino_acid_loop (:142-159)
    int c = (int)rng.below(MD_ALPHABET_SIZE);
    // push_modification (:182-208): the residue that was last loses its C-terminus modification ...
    if (!st->aa.empty() && (st->mod.back() & 1) && M.fix_pos[st->aa.back()] == 'C') { st->w -= M.fix[st->aa.back()]; st->mod.back() &= (uint8_t)~1; }
    st->aa.push_back(c); st->raw.push_back((uint8_t)kAlphabet[c]); st->mod.push_back(0);
    st->w += residue_mass((uint8_t)kAlphabet[c]);
    // ... and the new one takes its letter's fixed modification: anywhere, C-terminus (it is the last one now), N-terminus if it is the first
    if (M.fix_applies(c, (uint32_t)st->aa.size() - 1, (uint32_t)st->aa.size())) { st->w += M.fix[c]; st->mod.back() = 1; }
    if (st->w > pr.hi) break;                           // GreaterThenMassTolerance (:294-300)
    if (st->aa.size() > MD_MAX_PEPTIDE_LEN) return false;
  }
  if (st->aa.size() > MD_MAX_PEPTIDE_LEN) return false;
  const uint32_t L = (uint32_t)st->aa.size();
  for (int t = 0; t < 100; t++) {                       // 'tries (:453)
    g_stat_tries++;
    for (uint32_t i = 0; i < L; i++) {                  // 'sequence (:454)
      g_stat_evals++;
      int cur = st->aa[i];
      int64_t best = std::llabs(pr.mass - st->w); int bestc = cur;
      for (int c = 0; c < MD_ALPHABET_SIZE; c++) {      // 'swaps (:459-468), alphabet order
        if (c == cur) continue;
        int64_t d = std::llabs(pr.mass - (st->w + delta[cur * MD_ALPHABET_SIZE + c]));
        if (d < best) { best = d; bestc = c; }
      }
      if (bestc != cur) {
        replace_at(M, st, i, bestc);
        if (in_window(st->w, pr.lo, pr.hi)) return true;          // :483
        if (try_variable(M, st, pr.lo, pr.hi)) return true;       // :484
      }
    }
    // random kick (:489-505): position and letter are both uniform draws; here they come from ONE word r of the
    // attempt's stream (position = high half of r*L, letter = high half of low32(r*L)*21)
    const uint64_t rl = (uint64_t)rng.next() * L;
    uint32_t i = (uint32_t)(rl >> 32);
    int c = (int)(((uint64_t)(uint32_t)rl * MD_ALPHABET_SIZE) >> 32);
    replace_at(M, st, i, c);
  }
  return false;
}

void substitution_map(const ModSet& M, int64_t* out) {  // decoy_generator.rs:301-324
  for (int a = 0; a < MD_ALPHABET_SIZE; a++)
    for (int b = 0; b < MD_ALPHABET_SIZE; b++) out[a * MD_ALPHABET_SIZE + b] = M.mprime(b) - M.mprime(a);
}

uint32_t attempt_cap(uint32_t n) { return 16u * n + 1024u; }


// Stored decoys first (tasks/identification.rs:259-283): the same window/count queries as the targets
// (Decoy::find_where), the same ModifiedPeptide filter (from_decoy + try_variable_modifications), until n decoys.
// The reference takes them in database row order; here: in store-index order (W*, then canonical store order).
void decoys_reuse(const md_ctx* ctx, const md_precursor& pr, uint32_t n, std::vector<Decoy>* out) {
  const DecoyStore& D = ctx->store; const ModSet& M = ctx->mods;
  if (!D.indexed || D.seq.empty()) return;
  size_t b = std::lower_bound(D.key.begin(), D.key.end(), pr.lo) - D.key.begin();
  size_t e = std::upper_bound(D.key.begin(), D.key.end(), pr.hi) - D.key.begin();
  ModState st;
  for (size_t i = b; i < e && out->size() < n; i++) {
    uint32_t d = D.ord[i];
    if (!fanout_admits(M, &D.counts[(size_t)d * MD_ALPHABET_SIZE], pr.mass, pr.lo)) continue;
    const std::string& q = D.seq[d];
    from_string(M, (const uint8_t*)q.data(), (uint32_t)q.size(), &st);
    bool ok = in_window(st.w, pr.lo, pr.hi);
    if (!ok) ok = try_variable(M, &st, pr.lo, pr.hi);
    if (ok) out->push_back({q, var_mask_of(st), D.weight[d], st.w, MD_DECOY_STORED});
  }
}

void decoys_random(const md_ctx* ctx, const md_precursor& pr, uint32_t n, uint64_t seed,
                   std::vector<Decoy>* out) {
  int64_t delta[MD_ALPHABET_SIZE * MD_ALPHABET_SIZE];
  substitution_map(ctx->mods, delta);
  decoys_reuse(ctx, pr, n, out);
  std::unordered_set<std::string> seen;
  for (const Decoy& d : *out) seen.insert(d.seq);       // a generated decoy equal to a reused one is a duplicate
  ModState st;
  const uint32_t cap = attempt_cap(n);
  for (uint32_t a = 0; a < cap && out->size() < n; a++) {
    g_stat_attempts++;
    if (!random_attempt(ctx->mods, pr, seed, a, delta, &st)) continue;
    g_stat_success++;
    std::string s((const char*)st.raw.data(), st.raw.size());
    if (ctx->peps.by_seq.count(s)) continue;             // Decoy::is_peptide (decoy.rs:49-60)
    if (!seen.insert(s).second) continue;                // HashSet<Decoy> (decoy_generator.rs:40,164)
    out->push_back({s, var_mask_of(st), sequence_weight(st.raw.data(), (uint32_t)st.raw.size()), st.w, a});
  }
}

// vary_targets (decoy_generator.rs:265-296), made counter-based: attempt a shuffles target
// (a mod T) of the spectrum (Fisher-Yates on the original sequence, keyed by a); accepted when
// the fixed-mod weight is in the window, it is not a peptide and not seen before.
void decoys_permute(const md_ctx* ctx, const md_precursor& pr, uint32_t n, uint64_t seed,
                    std::vector<Decoy>* out) {
  std::vector<Candidate> targets;
  candidates_for(ctx, pr, &targets);
  const uint32_t T = (uint32_t)targets.size();
  if (!T) return;
  std::unordered_set<std::string> seen;
  ModState st;
  const uint64_t cap = std::min<uint64_t>((uint64_t)T * 1000u, attempt_cap(n));
  for (uint32_t a = 0; a < cap && out->size() < n; a++) {
    const std::string& src = ctx->peps.seq[targets[a % T].pep];
    std::string s = src;
    Philox rng(seed, pr.spectrum_id, a, kTagPermute);
    for (uint32_t i = (uint32_t)s.size(); i > 1; i--) { uint32_t j = rng.below(i); std::swap(s[i - 1], s[j]); }
    from_string(ctx->mods, (const uint8_t*)s.data(), (uint32_t)s.size(), &st);
    if (!in_window(st.w, pr.lo, pr.hi)) continue;
    if (ctx->peps.by_seq.count(s)) continue;
    if (!seen.insert(s).second) continue;
    out->push_back({s, 0, sequence_weight((const uint8_t*)s.data(), (uint32_t)s.size()), st.w, a});
  }
}

// EXHAUSTIVE (builder-defined): compositions (count vectors over MD_ALPHABET, fixed mods folded
// into the letter masses, no variable mods) with H2O + sum in [lo,hi] and 1 <= length <= 60, in
// ascending lexicographic order of the count vector (c_A, c_R, ..., c_Y); for each composition its
// distinct permutations in ascending lexicographic order of alphabet indices; peptides skipped;
// stop after n.  `attempt` = ordinal of the emitted sequence in that enumeration (peptides counted).
struct ExhaustiveEnum {
  const md_ctx* ctx; const md_precursor* pr; uint32_t n; std::vector<Decoy>* out;
  int64_t m[MD_ALPHABET_SIZE]; int64_t min_suffix[MD_ALPHABET_SIZE + 1], max_suffix[MD_ALPHABET_SIZE + 1];
  int cnt[MD_ALPHABET_SIZE]; uint32_t ordinal = 0;
  bool emit_perms() {
    std::vector<int> cur;
    for (int a = 0; a < MD_ALPHABET_SIZE; a++) for (int k = 0; k < cnt[a]; k++) cur.push_back(a);
    do {
      std::string s; for (int a : cur) s.push_back(kAlphabet[a]);
      uint32_t ord = ordinal++;
      if (!ctx->peps.by_seq.count(s)) {
        int64_t wu = sequence_weight((const uint8_t*)s.data(), (uint32_t)s.size());
        int64_t w = convert_mass_to_int(18.010565); for (int a : cur) w += m[a];
        out->push_back({s, 0, wu, w, ord});
        if (out->size() >= n) return true;
      }
    } while (std::next_permutation(cur.begin(), cur.end()));
    return false;
  }
  // remaining window [rlo, rhi] for letters a..20, `len` residues so far
  bool rec(int a, int64_t rlo, int64_t rhi, int len) {
    if (a == MD_ALPHABET_SIZE - 1) {
      // last letter: count fixed by the remainder
      int64_t ma = m[a];
      int64_t kmin = rlo <= 0 ? 0 : (rlo + ma - 1) / ma, kmax = rhi / ma;
      for (int64_t k = kmin; k <= kmax && len + k <= MD_MAX_PEPTIDE_LEN; k++) {
        if (len + k == 0) continue;
        cnt[a] = (int)k;
        if (emit_perms()) return true;
      }
      cnt[a] = 0;
      return false;
    }
    for (int k = 0; len + k <= MD_MAX_PEPTIDE_LEN; k++) {
      int64_t used = (int64_t)k * m[a];
      if (used > rhi) break;
      cnt[a] = k;
      if (rec(a + 1, rlo - used, rhi - used, len + k)) return true;
    }
    cnt[a] = 0;
    return false;
  }
};

void decoys_exhaustive(const md_ctx* ctx, const md_precursor& pr, uint32_t n, std::vector<Decoy>* out) {
  ExhaustiveEnum E; E.ctx = ctx; E.pr = &pr; E.n = n; E.out = out;
  for (int a = 0; a < MD_ALPHABET_SIZE; a++) { E.m[a] = ctx->mods.mprime(a); E.cnt[a] = 0; }
  const int64_t h2o = convert_mass_to_int(18.010565);
  if (n == 0 || pr.hi < h2o) return;
  E.rec(0, pr.lo - h2o, pr.hi - h2o, 0);
}

// ---------------------------------------------------------------------------------------
// scoring (builder-defined Comet-style fast xcorr in exact integers; see header comment)
// ---------------------------------------------------------------------------------------
const int kXcorrOffset = 75;      // Comet iXcorrProcessingOffset
const int kQ = 16;                // fixed-point fraction bits of the normalised intensities

struct BinnedSpectrum {
  bool scored = false;
  int64_t w = 0;                       // bin width uDa
  std::vector<int32_t> bin, yq;        // sorted unique bins, quantised normalised intensity
  std::vector<int64_t> prefix;         // prefix sums of yq
};

void bin_spectrum(const md_spectra* S, uint32_t s, int64_t P, int64_t w, uint32_t min_peaks, BinnedSpectrum* B) {
  B->scored = false; B->w = w; B->bin.clear(); B->yq.clear(); B->prefix.clear();
  std::vector<int32_t> bins; std::vector<double> raw;
  double gmax = 0; int32_t hbin = 0;
  for (uint64_t i = S->peak_off[s]; i < S->peak_off[s + 1]; i++) {
    double mz = S->peak_mz[i]; float I = S->peak_intensity[i];
    if (!(I > 0.0f) || !(mz > 0.0) || !(mz < 1.0e7)) continue;
    int64_t mzint = (int64_t)(mz * 1000000.0);
    if (!(mzint > 0) || !(mzint < P + 50000000LL)) continue;   // Comet: ion < ExpPepMass + 50
    int32_t b = (int32_t)(mzint / w) + 1;
    double r = std::sqrt((double)I);
    bins.push_back(b); raw.push_back(r);
    if (r > gmax) gmax = r;
    if (b > hbin) hbin = b;
  }
  if (bins.size() < min_peaks || bins.empty()) return;
  const int32_t wsize = hbin / 10 + 1;
  double winmax[10] = {0};
  for (size_t i = 0; i < bins.size(); i++) { int k = bins[i] / wsize; if (raw[i] > winmax[k]) winmax[k] = raw[i]; }
  const double thr = 0.05 * gmax;
  std::map<int32_t, int32_t> dense;
  for (size_t i = 0; i < bins.size(); i++) {
    if (!(raw[i] > thr)) continue;
    double scale = 50.0 / winmax[bins[i] / wsize];
    double y = raw[i] * scale;
    int32_t q = (int32_t)(y * 65536.0 + 0.5);
    auto it = dense.find(bins[i]);
    if (it == dense.end()) dense[bins[i]] = q; else if (q > it->second) it->second = q;
  }
  int64_t acc = 0;
  for (auto& kv : dense) { B->bin.push_back(kv.first); B->yq.push_back(kv.second); acc += kv.second; B->prefix.push_back(acc); }
  B->scored = true;
}

// T[b] = 151*yq[b] - sum_{j=b-75..b+75} yq[j]  ( = 150 * 2^16 * fast_xcorr[b] )
int64_t table_at(const BinnedSpectrum& B, int64_t b) {
  if (B.bin.empty()) return 0;
  auto cum = [&](int64_t x) -> int64_t {  // sum of yq over bins <= x
    size_t k = std::upper_bound(B.bin.begin(), B.bin.end(), (int32_t)std::min<int64_t>(x, INT32_MAX)) - B.bin.begin();
    if (x < 0) k = 0;
    return k ? B.prefix[k - 1] : 0;
  };
  int64_t win = cum(b + kXcorrOffset) - cum(b - kXcorrOffset - 1);
  int64_t here = cum(b) - cum(b - 1);
  return 151 * here - win;
}

int64_t score_candidate(const ModSet& M, const BinnedSpectrum& B, const uint8_t* seq, uint32_t len, uint64_t mask,
                        uint32_t z, uint32_t max_frag_charge) {
  if (!B.scored || len < 2) return 0;
  uint32_t nch = z > 1 ? z - 1 : 1;
  if (nch > max_frag_charge) nch = max_frag_charge;
  if (nch < 1) nch = 1;
  int64_t m[MD_MAX_PEPTIDE_LEN + 4]; int64_t total = 0;
  for (uint32_t i = 0; i < len; i++) {
    int a = alpha_index(seq[i]);
    int64_t v = residue_mass(seq[i]);
    if (a >= 0) {
      if (M.fix_applies(a, i, len)) v += M.fix[a];   // (a terminal modification adds to the end residue's mass: every b ion holds the first residue, every y ion the last)
      if ((mask >> i) & 1) v += M.var[a];
    }
    m[i] = v; total += v;
  }
  int64_t raw = 0, bsum = 0;
  for (uint32_t k = 1; k < len; k++) {
    bsum += m[k - 1];
    int64_t ysum = total - bsum + MD_WATER_UDA;
    for (uint32_t c = 1; c <= nch; c++) {
      int64_t bb = (bsum + (int64_t)c * MD_PROTON_UDA) / ((int64_t)c * B.w) + 1;
      int64_t yb = (ysum + (int64_t)c * MD_PROTON_UDA) / ((int64_t)c * B.w) + 1;
      raw += table_at(B, bb) + table_at(B, yb);
    }
  }
  return raw;
}

inline float final_score(int64_t raw) { return (float)(0.005 * (double)raw / (150.0 * 65536.0)); }

int validate_spectra(md_ctx* ctx, const md_spectra* S) {
  if (!S || (S->n && (!S->precursor_mz || !S->charge || !S->peak_off))) return fail(ctx, MD_ERR_INVALID, "spectra: null array");
  for (uint32_t s = 0; s < S->n; s++) {
    if (S->peak_off[s + 1] < S->peak_off[s]) return fail(ctx, MD_ERR_INVALID, "spectra: peak_off not monotone");
    for (uint64_t i = S->peak_off[s] + 1; i < S->peak_off[s + 1]; i++)
      if (S->peak_mz[i] < S->peak_mz[i - 1]) return fail(ctx, MD_ERR_INVALID, "spectra: peaks of a spectrum must be sorted by m/z");
    if (S->charge[s] == 0) return fail(ctx, MD_ERR_INVALID, "spectra: charge 0");
    if (!(std::isfinite(S->precursor_mz[s]) && S->precursor_mz[s] > 0.0 && S->precursor_mz[s] < 1.0e7)) return fail(ctx, MD_ERR_INVALID, "spectra: precursor m/z must be finite and in (0, 1e7)");
  }
  return MD_OK;
}

int make_precursor(const md_spectra* S, const md_search_params* p, uint32_t s, md_precursor* pr) {
  int64_t P, lo, hi;
  precursor_window(S->precursor_mz[s], S->charge[s], p->lower_ppm, p->upper_ppm, &P, &lo, &hi);
  if (p->abs_lower_uda != 0 || p->abs_upper_uda != 0) { lo = P - p->abs_lower_uda; hi = P + p->abs_upper_uda; }
  pr->mass = P; pr->lo = lo; pr->hi = hi; pr->charge = S->charge[s];
  pr->spectrum_id = S->spectrum_id ? S->spectrum_id[s] : s;
  return MD_OK;
}

int gen_decoys(md_ctx* ctx, const md_precursor& pr, uint32_t n, int mode, uint64_t seed, std::vector<Decoy>* out) {
  switch (mode) {
    case MD_DECOY_REFERENCE_RANDOM: decoys_random(ctx, pr, n, seed, out); return MD_OK;
    case MD_DECOY_PERMUTE_TARGET: decoys_permute(ctx, pr, n, seed, out); return MD_OK;
    case MD_DECOY_EXHAUSTIVE: decoys_exhaustive(ctx, pr, n, out); return MD_OK;
    default: return MD_ERR_INVALID;
  }
}

template <class F>
void parallel_for(uint32_t n, uint32_t n_threads, F f) {
  if (n_threads <= 1 || n < 2) { for (uint32_t i = 0; i < n; i++) f(i); return; }
  std::atomic<uint32_t> next(0);
  std::vector<std::thread> th;
  for (uint32_t t = 0; t < n_threads; t++)
    th.emplace_back([&]() { for (;;) { uint32_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
  for (auto& t : th) t.join();
}

template <class T> T* dup(const std::vector<T>& v) {
  T* p = (T*)std::malloc(std::max<size_t>(1, v.size()) * sizeof(T));
  if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

int fill_decoy_table(const std::vector<std::vector<Decoy>>& per, md_decoy_table* out) {
  std::vector<uint64_t> off(1, 0), seq_off(1, 0), mask; std::vector<uint8_t> seq; std::vector<int64_t> wt, mw; std::vector<uint32_t> att;
  for (auto& v : per) {
    for (auto& d : v) {
      seq.insert(seq.end(), d.seq.begin(), d.seq.end()); seq_off.push_back(seq.size());
      mask.push_back(d.mask); wt.push_back(d.weight); mw.push_back(d.w); att.push_back(d.attempt);
    }
    off.push_back(mask.size());
  }
  out->n_spectra = (uint32_t)per.size(); out->n = mask.size(); out->seq_bytes = seq.size();
  out->off = dup(off); out->seq = dup(seq); out->seq_off = dup(seq_off); out->var_mask = dup(mask);
  out->weight = dup(wt); out->mod_weight = dup(mw); out->attempt = dup(att);
  return MD_OK;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

MD_API void md_oracle_decoy_stats(uint64_t out[4], int reset) {
  out[0] = g_stat_attempts; out[1] = g_stat_success; out[2] = g_stat_tries; out[3] = g_stat_evals;
  if (reset) { g_stat_attempts = 0; g_stat_success = 0; g_stat_tries = 0; g_stat_evals = 0; }
}

const char* md_backend_name(void) { return "cpu-oracle"; }

int md_create(const md_config* cfg, md_ctx** out) {
  if (!out) return fail(nullptr, MD_ERR_INVALID, "md_create: out is NULL");
  md_ctx* c = new md_ctx();
  c->n_threads = (cfg && cfg->n_threads) ? cfg->n_threads : 1;
  *out = c;
  return MD_OK;
}
void md_destroy(md_ctx* ctx) { delete ctx; }
const char* md_last_error(const md_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }
void md_free(void* p) { std::free(p); }
int md_sync(md_ctx*) { return MD_OK; }
// multi-GPU entry points: the checker is one process on the host; a single rank gathers by copying
int md_comm_unique_id(uint8_t* id) { if (!id) return MD_ERR_INVALID; memset(id, 0, MD_COMM_ID_BYTES); return MD_OK; }
int md_comm_init(md_ctx* ctx, int32_t rank, int32_t nranks, const uint8_t*) {
  if (!ctx || rank != 0) return MD_ERR_INVALID;
  return nranks == 1 ? MD_OK : MD_ERR_UNSUPPORTED;
}
int md_gather_psms(md_ctx* ctx, const md_psm* local, uint64_t rows, md_psm* all) {
  if (!ctx || (rows && (!local || !all))) return MD_ERR_INVALID;
  if (rows && local != all) memmove(all, local, rows * sizeof(md_psm));
  return MD_OK;
}
int md_comm_destroy(md_ctx* ctx) { return ctx ? MD_OK : MD_ERR_INVALID; }
void* md_stream_handle(md_ctx*) { return nullptr; }

int64_t md_residue_mass(uint8_t c) { return residue_mass(c); }
int64_t md_sequence_weight(const uint8_t* seq, uint32_t len) { return sequence_weight(seq, len); }
int md_precursor_window(double mz, uint32_t charge, int64_t lppm, int64_t uppm, int64_t* P, int64_t* lo, int64_t* hi) {
  if (!P || !lo || !hi || charge == 0 || charge > 255) return fail(nullptr, MD_ERR_INVALID, "md_precursor_window: bad argument");
  precursor_window(mz, charge, lppm, uppm, P, lo, hi);
  return MD_OK;
}

int md_set_modifications(md_ctx* ctx, const md_modification* mods, uint32_t n, uint32_t max_var) {
  if (!ctx || (n && !mods)) return fail(ctx, MD_ERR_INVALID, "md_set_modifications: null argument");
  if (max_var > 255) return fail(ctx, MD_ERR_INVALID, "md_set_modifications: max_variable_mods > 255 (u8 in the reference)");
  ModSet M; M.nvar = max_var;
  for (uint32_t i = 0; i < n; i++) {
    uint8_t pos = (uint8_t)std::toupper(mods[i].position);
    if (pos != 'A' && pos != 'N' && pos != 'C') return fail(ctx, MD_ERR_INVALID, "modification position must be A, N or C");  // modification.rs:24-33
    uint8_t aa = (uint8_t)std::toupper(mods[i].amino_acid);
    int a = alpha_index(aa);
    if (a < 0) return fail(ctx, MD_ERR_INVALID, "modification on a letter without a <x>_count column (alphabet " MD_ALPHABET ")");
    if (mods[i].is_fix) { M.has_fix[a] = true; M.fix[a] = mods[i].mono_mass; M.fix_pos[a] = pos; }
    else { M.has_var[a] = true; M.var[a] = mods[i].mono_mass; M.var_pos[a] = pos; }
  }
  M.has_terminal = false;   // (of the modifications that survived: a later entry of a letter replaces an earlier one)
  for (int a = 0; a < MD_ALPHABET_SIZE; a++) if ((M.has_fix[a] && M.fix_pos[a] != 'A') || (M.has_var[a] && M.var_pos[a] != 'A')) M.has_terminal = true;
  for (int a = 0; a < MD_ALPHABET_SIZE; a++)
    if (M.has_fix[a] || M.has_var[a]) {
      if (residue_mass((uint8_t)kAlphabet[a]) + M.merged(a) <= 0) return fail(ctx, MD_ERR_INVALID, "modified residue mass must be positive");
      if (M.mprime(a) <= 0) return fail(ctx, MD_ERR_INVALID, "modified residue mass must be positive");
      M.letter_chars.push_back((uint8_t)kAlphabet[a]);
    }
  std::sort(M.letter_chars.begin(), M.letter_chars.end());  // identification.rs:177-178 (sort by char)
  for (uint8_t c : M.letter_chars) M.letters.push_back(alpha_index(c));
  M.set = true;
  ctx->mods = M;
  ctx->index.ready = false; ctx->store.indexed = false;
  return MD_OK;
}

int md_substitution_map(md_ctx* ctx, int64_t* out) {
  if (!ctx || !out) return fail(ctx, MD_ERR_INVALID, "md_substitution_map: null argument");
  substitution_map(ctx->mods, out);
  return MD_OK;
}

int md_digest(md_ctx* ctx, const uint8_t* residues, const uint64_t* off, uint32_t n_prot, const md_digest_params* p, uint64_t* n_out) {
  if (!ctx || !p || (n_prot && (!residues || !off))) return fail(ctx, MD_ERR_INVALID, "md_digest: null argument");
  if (p->max_len > MD_MAX_PEPTIDE_LEN) return fail(ctx, MD_ERR_INVALID, "md_digest: max_len > 60 (tasks/digestion.rs:101)");
  if (p->max_missed_cleavages > 60) return fail(ctx, MD_ERR_INVALID, "md_digest: max_missed_cleavages > 60 (tasks/digestion.rs:74)");
  if (p->min_len < 1 || p->min_len > p->max_len) return fail(ctx, MD_ERR_INVALID, "md_digest: need 1 <= min_len <= max_len");
  for (uint32_t i = 0; i < n_prot; i++) if (off[i + 1] < off[i]) return fail(ctx, MD_ERR_INVALID, "md_digest: protein_offsets not monotone");
  std::vector<Occ> occ;
  std::vector<std::pair<size_t, size_t>> pieces;
  for (uint32_t pr = 0; pr < n_prot; pr++) {
    const uint8_t* s = residues + off[pr]; size_t n = off[pr + 1] - off[pr];
    pieces.clear(); split_trypsin(s, n, &pieces);
    for (size_t i = 0; i < pieces.size(); i++) {
      for (uint32_t mc = 0; mc <= p->max_missed_cleavages; mc++) {   // digest_enzym.rs:64-86
        size_t j = i + mc; if (j >= pieces.size()) break;
        size_t b = pieces[i].first, e = pieces[j].second, len = e - b;
        if (len > p->max_len) break;                                  // only grows from here
        if (len >= p->min_len) {
          std::string g((const char*)s + b, len);
          for (auto& ch : g) if (ch == 'I' || ch == 'L') ch = 'J';   // amino_acid.rs:139-141
          occ.push_back({std::move(g), (uint8_t)mc, pr});
        }
      }
    }
  }
  // unique by generalized sequence (peptide.rs:277-291; schema.sql:41); canonical order
  struct Rep { uint64_t first; int64_t w; uint64_t h; uint8_t mc; std::vector<uint32_t> prots; };
  std::unordered_map<std::string, size_t> seen; std::vector<Rep> reps; std::vector<const std::string*> rep_seq;
  for (size_t i = 0; i < occ.size(); i++) {
    auto it = seen.find(occ[i].g);
    if (it == seen.end()) {
      seen.emplace(occ[i].g, reps.size());
      const uint8_t* d = (const uint8_t*)occ[i].g.data();
      reps.push_back({i, sequence_weight(d, (uint32_t)occ[i].g.size()), hash64(d, (uint32_t)occ[i].g.size()), occ[i].mc, {occ[i].prot}});
      rep_seq.push_back(&occ[i].g);
    } else {
      Rep& r = reps[it->second];
      if (occ[i].mc < r.mc) r.mc = occ[i].mc;
      if (r.prots.back() != occ[i].prot) r.prots.push_back(occ[i].prot);  // protein ordinals arrive ascending
    }
  }
  std::vector<size_t> order(reps.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    if (reps[a].w != reps[b].w) return reps[a].w < reps[b].w;
    if (reps[a].h != reps[b].h) return reps[a].h < reps[b].h;
    return reps[a].first < reps[b].first;
  });
  Peptides P; P.seq.reserve(order.size());
  P.counts.assign(order.size() * MD_ALPHABET_SIZE, 0);
  for (size_t k = 0; k < order.size(); k++) {
    const Rep& r = reps[order[k]]; const std::string& s = *rep_seq[order[k]];
    P.seq.push_back(s); P.weight.push_back(r.w); P.mc.push_back(r.mc); P.assoc.push_back(r.prots);
    for (char ch : s) { int a = alpha_index((uint8_t)ch); if (a >= 0) P.counts[k * MD_ALPHABET_SIZE + a]++; }  // peptide_interface.rs:22-28
    P.by_seq.emplace(s, (uint32_t)k);
  }
  P.ready = true;
  ctx->peps = std::move(P);
  ctx->index.ready = false;
  if (n_out) *n_out = ctx->peps.seq.size();
  return MD_OK;
}

int md_peptides_export(md_ctx* ctx, md_peptide_table* out) {
  if (!ctx || !out) return fail(ctx, MD_ERR_INVALID, "md_peptides_export: null argument");
  if (!ctx->peps.ready) return fail(ctx, MD_ERR_STATE, "md_peptides_export: no digest yet");
  const Peptides& P = ctx->peps;
  std::vector<uint8_t> seq; std::vector<uint64_t> so(1, 0), ao(1, 0); std::vector<uint32_t> ap;
  for (size_t i = 0; i < P.seq.size(); i++) {
    seq.insert(seq.end(), P.seq[i].begin(), P.seq[i].end()); so.push_back(seq.size());
    ap.insert(ap.end(), P.assoc[i].begin(), P.assoc[i].end()); ao.push_back(ap.size());
  }
  out->n = P.seq.size(); out->seq_bytes = seq.size(); out->n_assoc = ap.size();
  out->seq = dup(seq); out->seq_off = dup(so); out->missed_cleavages = dup(P.mc); out->weight = dup(P.weight);
  out->counts = dup(P.counts); out->assoc_off = dup(ao); out->assoc_protein = dup(ap);
  return MD_OK;
}
void md_peptide_table_free(md_peptide_table* t) {
  if (!t) return;
  std::free(t->seq); std::free(t->seq_off); std::free(t->missed_cleavages); std::free(t->weight);
  std::free(t->counts); std::free(t->assoc_off); std::free(t->assoc_protein);
  std::memset(t, 0, sizeof(*t));
}

namespace {
void index_store(md_ctx* ctx) {
  DecoyStore& D = ctx->store; const ModSet& M = ctx->mods;
  const size_t n = D.seq.size();
  std::vector<std::pair<int64_t, uint32_t>> kv(n);
  for (size_t i = 0; i < n; i++) {
    int64_t k = D.weight[i];
    for (int a : M.letters) k += (int64_t)D.counts[i * MD_ALPHABET_SIZE + a] * M.merged(a);
    kv[i] = {k, (uint32_t)i};
  }
  std::sort(kv.begin(), kv.end());
  D.key.resize(n); D.ord.resize(n);
  for (size_t i = 0; i < n; i++) { D.key[i] = kv[i].first; D.ord[i] = kv[i].second; }
  D.indexed = true;
}
}  // namespace

int md_index_build(md_ctx* ctx) {
  if (!ctx) return fail(ctx, MD_ERR_INVALID, "md_index_build: null ctx");
  if (!ctx->peps.ready) return fail(ctx, MD_ERR_STATE, "md_index_build: md_digest first");
  if (!ctx->mods.set) return fail(ctx, MD_ERR_STATE, "md_index_build: md_set_modifications first");
  const Peptides& P = ctx->peps; const ModSet& M = ctx->mods;
  size_t n = P.seq.size();
  std::vector<std::pair<int64_t, uint32_t>> kv(n);
  for (size_t i = 0; i < n; i++) {
    int64_t k = P.weight[i];
    for (int a : M.letters) k += (int64_t)P.counts[i * MD_ALPHABET_SIZE + a] * M.merged(a);
    kv[i] = {k, (uint32_t)i};
  }
  std::sort(kv.begin(), kv.end());
  ctx->index.key.resize(n); ctx->index.pep.resize(n);
  for (size_t i = 0; i < n; i++) { ctx->index.key[i] = kv[i].first; ctx->index.pep[i] = kv[i].second; }
  ctx->fixed.clear();
  if (ctx->var_mode == MD_VARMOD_EXPANDED) {
    ctx->fixed.resize(n);
    for (size_t i = 0; i < n; i++) {
      const uint32_t p = ctx->index.pep[i];
      ModState st;
      from_string(M, (const uint8_t*)P.seq[p].data(), (uint32_t)P.seq[p].size(), &st);   // weight with the fixed modifications where their positions allow
      ctx->fixed[i] = {st.w, (uint32_t)i};
    }
    std::sort(ctx->fixed.begin(), ctx->fixed.end());
  }
  ctx->index.ready = true;
  index_store(ctx);
  return MD_OK;
}

int md_set_variable_mode(md_ctx* ctx, int mode) {
  if (!ctx) return fail(ctx, MD_ERR_INVALID, "md_set_variable_mode: null ctx");
  if (mode != MD_VARMOD_REFERENCE && mode != MD_VARMOD_EXPANDED) return fail(ctx, MD_ERR_INVALID, "md_set_variable_mode: unknown mode");
  if (mode != ctx->var_mode) { ctx->var_mode = mode; ctx->index.ready = false; ctx->store.indexed = false; }
  return MD_OK;
}

int md_decoy_store_set(md_ctx* ctx, const uint8_t* seq, const uint64_t* off, uint64_t n) {
  if (!ctx || (n && (!seq || !off))) return fail(ctx, MD_ERR_INVALID, "md_decoy_store_set: null argument");
  struct E { int64_t w; uint64_t h; std::string s; };
  std::vector<E> v; v.reserve(n);
  for (uint64_t i = 0; i < n; i++) {
    if (off[i + 1] < off[i]) return fail(ctx, MD_ERR_INVALID, "md_decoy_store_set: offsets not monotone");
    const uint64_t L = off[i + 1] - off[i];
    if (L == 0 || L > MD_MAX_PEPTIDE_LEN) return fail(ctx, MD_ERR_INVALID, "md_decoy_store_set: sequence length must be 1..60");
    std::string q((const char*)seq + off[i], (size_t)L);
    for (char c : q) if (alpha_index((uint8_t)c) < 0) return fail(ctx, MD_ERR_INVALID, "md_decoy_store_set: letter outside the decoy alphabet " MD_ALPHABET);
    v.push_back({sequence_weight((const uint8_t*)q.data(), (uint32_t)L), hash64((const uint8_t*)q.data(), (uint32_t)L), q});
  }
  std::sort(v.begin(), v.end(), [](const E& a, const E& b) { return a.w != b.w ? a.w < b.w : a.h != b.h ? a.h < b.h : a.s < b.s; });
  DecoyStore D;
  for (size_t i = 0; i < v.size(); i++) {
    if (i && v[i].s == v[i - 1].s) continue;              // UNIQUE (aa_sequence, weight)
    D.seq.push_back(v[i].s); D.weight.push_back(v[i].w);
    for (int a = 0; a < MD_ALPHABET_SIZE; a++) D.counts.push_back((int16_t)std::count(v[i].s.begin(), v[i].s.end(), kAlphabet[a]));
  }
  ctx->store = std::move(D);
  if (ctx->index.ready) index_store(ctx);
  return MD_OK;
}

int md_index_stats_get(md_ctx* ctx, md_index_stats* out) {
  if (!ctx || !out) return fail(ctx, MD_ERR_INVALID, "md_index_stats_get: null argument");
  if (!ctx->index.ready) return fail(ctx, MD_ERR_STATE, "md_index_stats_get: md_index_build first");
  std::memset(out, 0, sizeof(*out));
  out->n_peptides = ctx->index.key.size();
  for (auto& s : ctx->peps.seq) out->seq_bytes += s.size();
  if (!ctx->index.key.empty()) { out->min_key = ctx->index.key.front(); out->max_key = ctx->index.key.back(); }
  return MD_OK;
}

int md_window_search(md_ctx* ctx, const int64_t* lo, const int64_t* hi, uint32_t n, uint64_t* begin, uint64_t* end) {
  if (!ctx || (n && (!lo || !hi || !begin || !end))) return fail(ctx, MD_ERR_INVALID, "md_window_search: null argument");
  if (!ctx->index.ready) return fail(ctx, MD_ERR_STATE, "md_window_search: md_index_build first");
  const auto& K = ctx->index.key;
  for (uint32_t i = 0; i < n; i++) {
    begin[i] = std::lower_bound(K.begin(), K.end(), lo[i]) - K.begin();
    end[i] = std::upper_bound(K.begin(), K.end(), hi[i]) - K.begin();
    if (end[i] < begin[i]) end[i] = begin[i];
  }
  return MD_OK;
}

int md_index_export(md_ctx* ctx, uint64_t begin, uint64_t count, uint64_t* pid, int64_t* key) {
  if (!ctx) return fail(ctx, MD_ERR_INVALID, "md_index_export: null ctx");
  if (!ctx->index.ready) return fail(ctx, MD_ERR_STATE, "md_index_export: md_index_build first");
  if (begin + count > ctx->index.key.size()) return fail(ctx, MD_ERR_INVALID, "md_index_export: range out of bounds");
  for (uint64_t i = 0; i < count; i++) { if (pid) pid[i] = (uint64_t)ctx->index.pep[begin + i] + 1; if (key) key[i] = ctx->index.key[begin + i]; }
  return MD_OK;
}

int md_candidates(md_ctx* ctx, const md_precursor* pr, uint32_t n, md_candidate_table* out) {
  if (!ctx || !out || (n && !pr)) return fail(ctx, MD_ERR_INVALID, "md_candidates: null argument");
  if (!ctx->index.ready) return fail(ctx, MD_ERR_STATE, "md_candidates: md_index_build first");
  std::vector<std::vector<Candidate>> per(n);
  parallel_for(n, ctx->n_threads, [&](uint32_t s) { candidates_for(ctx, pr[s], &per[s]); });
  std::vector<uint64_t> off(1, 0), pid, mask; std::vector<int64_t> mw;
  for (auto& v : per) { for (auto& c : v) { pid.push_back((uint64_t)c.pep + 1); mask.push_back(c.mask); mw.push_back(c.w); } off.push_back(pid.size()); }
  out->n_spectra = n; out->n = pid.size(); out->off = dup(off); out->peptide_id = dup(pid); out->var_mask = dup(mask); out->mod_weight = dup(mw);
  return MD_OK;
}
void md_candidate_table_free(md_candidate_table* t) {
  if (!t) return;
  std::free(t->off); std::free(t->peptide_id); std::free(t->var_mask); std::free(t->mod_weight);
  std::memset(t, 0, sizeof(*t));
}

int md_generate_decoys(md_ctx* ctx, const md_precursor* pr, uint32_t n_spec, uint32_t n_per, int mode, uint64_t seed, md_decoy_table* out) {
  if (!ctx || !out || (n_spec && !pr)) return fail(ctx, MD_ERR_INVALID, "md_generate_decoys: null argument");
  if (!ctx->index.ready) return fail(ctx, MD_ERR_STATE, "md_generate_decoys: md_index_build first");
  if (mode < 0 || mode > 2) return fail(ctx, MD_ERR_INVALID, "md_generate_decoys: unknown mode");
  if (mode == MD_DECOY_EXHAUSTIVE && ctx->mods.has_terminal && n_spec && n_per) return fail(ctx, MD_ERR_UNSUPPORTED, "exhaustive decoys enumerate compositions: not defined with terminal modifications (the weight depends on the order)");
  std::vector<std::vector<Decoy>> per(n_spec);
  parallel_for(n_spec, ctx->n_threads, [&](uint32_t s) { gen_decoys(ctx, pr[s], n_per, mode, seed, &per[s]); });
  return fill_decoy_table(per, out);
}
void md_decoy_table_free(md_decoy_table* t) {
  if (!t) return;
  std::free(t->off); std::free(t->seq); std::free(t->seq_off); std::free(t->var_mask); std::free(t->weight); std::free(t->mod_weight); std::free(t->attempt);
  std::memset(t, 0, sizeof(*t));
}

int md_identify(md_ctx* ctx, const md_spectra* S, const md_search_params* p, md_psm* psms, md_identify_stats* stats, int64_t** all_scores, uint64_t** all_off) {
  if (!ctx || !S || !p || (S->n && p->top_k && !psms)) return fail(ctx, MD_ERR_INVALID, "md_identify: null argument");
  if (!ctx->index.ready) return fail(ctx, MD_ERR_STATE, "md_identify: md_index_build first");
  int rc = validate_spectra(ctx, S); if (rc) return rc;
  if (p->decoy_mode < 0 || p->decoy_mode > 2) return fail(ctx, MD_ERR_INVALID, "md_identify: unknown decoy mode");
  if (p->decoy_mode == MD_DECOY_EXHAUSTIVE && ctx->mods.has_terminal && S->n && p->n_decoys) return fail(ctx, MD_ERR_UNSUPPORTED, "exhaustive decoys enumerate compositions: not defined with terminal modifications (the weight depends on the order)");
  const int64_t w = (int64_t)std::llround(p->fragment_tolerance * 1000000.0);
  if (w < 100 || w > 2000000) return fail(ctx, MD_ERR_INVALID, "md_identify: fragment_tolerance must be in [0.0001, 2] Da");
  const uint32_t mfc = p->max_fragment_charge ? p->max_fragment_charge : 3;
  const uint32_t n = S->n, K = p->top_k;
  std::vector<std::vector<int64_t>> scores(n);
  std::vector<std::vector<Decoy>> decoys(n);
  std::vector<uint64_t> nt(n, 0);
  std::atomic<uint64_t> t_lookup(0), t_decoy(0), t_score(0);
  auto t0 = std::chrono::steady_clock::now();
  parallel_for(n, ctx->n_threads, [&](uint32_t s) {
    using clk = std::chrono::steady_clock;
    md_precursor pr; make_precursor(S, p, s, &pr);
    auto a = clk::now();
    std::vector<Candidate> targets; candidates_for(ctx, pr, &targets);
    auto b = clk::now();
    std::vector<Decoy>& dec = decoys[s];
    if (p->n_decoys) gen_decoys(ctx, pr, p->n_decoys, p->decoy_mode, p->seed, &dec);
    auto c = clk::now();
    BinnedSpectrum B; bin_spectrum(S, s, pr.mass, w, p->min_peaks, &B);
    std::vector<int64_t>& sc = scores[s]; sc.reserve(targets.size() + dec.size());
    for (auto& t : targets) { const std::string& q = ctx->peps.seq[t.pep]; sc.push_back(score_candidate(ctx->mods, B, (const uint8_t*)q.data(), (uint32_t)q.size(), t.mask, pr.charge, mfc)); }
    for (auto& d : dec) sc.push_back(score_candidate(ctx->mods, B, (const uint8_t*)d.seq.data(), (uint32_t)d.seq.size(), d.mask, pr.charge, mfc));
    auto e = clk::now();
    nt[s] = targets.size();
    // top-k rows
    std::vector<uint32_t> ord(sc.size());
    for (uint32_t i = 0; i < ord.size(); i++) ord[i] = i;
    uint32_t kk = std::min<uint32_t>(K, (uint32_t)ord.size());
    if (!B.scored) kk = 0;
    std::partial_sort(ord.begin(), ord.begin() + kk, ord.end(), [&](uint32_t x, uint32_t y) { return sc[x] != sc[y] ? sc[x] > sc[y] : x < y; });
    for (uint32_t r = 0; r < K; r++) {
      md_psm& row = psms[(size_t)s * K + r]; std::memset(&row, 0, sizeof(row));
      row.spectrum_id = pr.spectrum_id; row.charge = (uint8_t)pr.charge;
      row.n_targets = (uint32_t)targets.size(); row.n_decoys = (uint32_t)dec.size();
      if (r >= kk) continue;
      uint32_t i = ord[r];
      row.rank = (uint16_t)(r + 1);
      if (i < targets.size()) { row.is_decoy = 0; row.candidate = (uint64_t)targets[i].pep + 1; row.var_mask = targets[i].mask; row.mod_weight = targets[i].w; }
      else { const Decoy& d = dec[i - targets.size()]; row.is_decoy = 1; row.candidate = i - targets.size(); row.var_mask = d.mask; row.mod_weight = d.w; }
      row.raw_score = sc[i]; row.score = final_score(sc[i]);
    }
    t_lookup += std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count();
    t_decoy += std::chrono::duration_cast<std::chrono::nanoseconds>(c - b).count();
    t_score += std::chrono::duration_cast<std::chrono::nanoseconds>(e - c).count();
  });
  auto t1 = std::chrono::steady_clock::now();
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->n_spectra = n;
    for (uint32_t s = 0; s < n; s++) { stats->n_targets += nt[s]; stats->n_decoys += decoys[s].size(); if (decoys[s].size() < p->n_decoys) stats->n_less_decoys++; }
    double th = (double)std::max<uint32_t>(1, ctx->n_threads);
    stats->ms_lookup = t_lookup / 1e6 / th; stats->ms_decoys = t_decoy / 1e6 / th; stats->ms_score = t_score / 1e6 / th;
    stats->ms_total = std::chrono::duration<double, std::milli>(t1 - t0).count();
    stats->n_pairs = stats->n_targets + stats->n_decoys;
  }
  if (all_scores && all_off) {
    std::vector<uint64_t> off(1, 0); std::vector<int64_t> flat;
    for (auto& v : scores) { flat.insert(flat.end(), v.begin(), v.end()); off.push_back(flat.size()); }
    *all_scores = dup(flat); *all_off = dup(off);
  }
  if (p->keep_decoys) { ctx->last_decoys = std::move(decoys); ctx->have_last_decoys = true; }
  else ctx->have_last_decoys = false;
  return MD_OK;
}

int md_identify_device(md_ctx* ctx, const md_spectra*, const md_search_params*, md_psm*, md_identify_stats*) {
  return fail(ctx, MD_ERR_UNSUPPORTED, "md_identify_device: the CPU oracle has no device path");
}

int md_last_decoys_export(md_ctx* ctx, md_decoy_table* out) {
  if (!ctx || !out) return fail(ctx, MD_ERR_INVALID, "md_last_decoys_export: null argument");
  if (!ctx->have_last_decoys) return fail(ctx, MD_ERR_STATE, "md_last_decoys_export: no identify call with keep_decoys");
  return fill_decoy_table(ctx->last_decoys, out);
}

}  // extern "C"
