"""Per-spectrum drop-in outputs of `identification`: `<spectrum>.fasta`, `<spectrum>.comet.params`
and `<spectrum>.less_decoys`, byte-compatible with what the reference writes (tasks/identification.rs:323-368), so that
users can still run Comet on the GPU-generated candidate sets -- the only way to validate scores against the
reference pipeline, which has no scorer of its own.

Formats:
  target header  >PEPTIDE_<seq> MaxDecoyId=<id>[ ModRes=<summary>]      (models/peptides/peptide.rs:11,64-70,238-240)
  decoy header   >DECOY_<seq>[ ModRes=<summary>]                        (models/peptides/decoy.rs:10,76-82)
  summary        "(<count>|<accession>|<name>)" per key "<accession>|<name>", keys ascending
                                                                         (models/peptides/modified_peptide.rs:606-659)
  entry          header "\\n" sequence "\\n", deduplicated by sequence    (models/fasta_entry.rs:16-18,30-44)
"""
import os

AMINO_ACID_NAMES = {  # models/amino_acids/amino_acid.rs:7-35, column 1
    "A": "Alanine", "B": "Asparagine or aspartic acid", "R": "Arginine", "N": "Asparagine", "D": "Aspartic acid",
    "C": "Cysteine", "E": "Glutamic acid", "Q": "Glutamine", "G": "Glycine", "H": "Histidine", "I": "Isoleucine",
    "L": "Leucine", "J": "Isoleucine or Leucine", "K": "Lysine", "M": "Methionine", "F": "Phenylalanine", "P": "Proline",
    "O": "Pyrrolysine", "S": "Serine", "T": "Threonine", "U": "Selenocysteine", "V": "Valine", "W": "Tryptophan",
    "X": "Unknown Amino Acid", "Y": "Tyrosine", "Z": "Glutamine or glutamic acid"}
J_MONO_INT = 113084060


def rust_f64(x):
    """Rust's `{}` for an f64: shortest round-trip digits, no trailing `.0`, never scientific notation."""
    x = float(x)
    if x == int(x) and abs(x) < 1e16:
        return str(int(x))
    r = repr(x)
    if "e" in r or "E" in r:
        r = format(x, "f").rstrip("0")
    return r


def modification_summary(sequence, mods, var_mask=0):
    """ModifiedPeptide::get_modification_summary_for_header (modified_peptide.rs:606-659): a residue carries its letter's
    fixed modification where the modification's position allows (anywhere; N / C: first / last residue only), and its
    variable modification iff bit i of var_mask (the library sets that bit only where the variable modification's own
    slot -- side chain or terminus -- is free)."""
    fix = {m.amino_acid: m for m in mods if m.is_fix}
    var = {m.amino_acid: m for m in mods if not m.is_fix}
    last = len(sequence) - 1
    counts = {}
    for i, c in enumerate(sequence):
        hits = []
        m = fix.get(c)
        if m is not None and (m.position == "A" or (m.position == "N" and i == 0) or (m.position == "C" and i == last)):
            hits.append(m)
        if (var_mask >> i) & 1 and c in var and not (hits and hits[0].position == var[c].position):
            hits.append(var[c])
        for m in hits:
            key = "%s|%s" % (m.accession, m.name)
            counts[key] = counts.get(key, 0) + 1
    return "".join("(%d|%s)" % (counts[k], k) for k in sorted(counts))


def peptide_header(sequence, peptide_id, summary=""):
    h = ">PEPTIDE_%s MaxDecoyId=%d" % (sequence, peptide_id)
    return h + (" ModRes=%s" % summary if summary else "")


def decoy_header(sequence, summary=""):
    h = ">DECOY_%s" % sequence
    return h + (" ModRes=%s" % summary if summary else "")


def fasta_entry(header, sequence):
    return "%s\n%s\n" % (header, sequence)


# The Comet parameter file the reference emits (utility/comet_parameter.rs:6-124), as data: blank-line separated groups
# of `key = value`; values are Comet's, fixed by the reference (b/y ions, monoisotopic masses, ppm units, ...).
_COMET_GROUPS_BEFORE = [
    [("decoy_search", "0"), ("peff_format", "0"), ("peff_obo", "")],
    [("num_threads", "0")],
    [("peptide_mass_units", "2"), ("mass_type_parent", "1"), ("mass_type_fragment", "1"), ("precursor_tolerance_type", "1"),
     ("isotope_error", "3")],
    [("search_enzyme_number", "1"), ("num_enzyme_termini", "2"), ("allowed_missed_cleavage", "2")],
    [("max_variable_mods_in_peptide", "5"), ("require_variable_mod", "0")],
    [("theoretical_fragment_ions", "1"), ("use_A_ions", "0"), ("use_B_ions", "1"), ("use_C_ions", "0"), ("use_X_ions", "0"),
     ("use_Y_ions", "1"), ("use_Z_ions", "0"), ("use_NL_ions", "0")],
    [("output_sqtstream", "0"), ("output_sqtfile", "0"), ("output_txtfile", "1"), ("output_pepxmlfile", "0"),
     ("output_percolatorfile", "0"), ("print_expect_score", "1"), ("show_fragment_ions", "0")],
    [("sample_enzyme_number", "1")],
    [("scan_range", "0 0"), ("precursor_charge", "0 0"), ("override_charge", "0"), ("ms_level", "2"), ("activation_method", "ALL")],
    [("digest_mass_range", "600.0 5000.0"), ("skip_researching", "1"), ("max_fragment_charge", "3"), ("max_precursor_charge", "6"),
     ("nucleotide_reading_frame", "0"), ("clip_nterm_methionine", "0"), ("spectrum_batch_size", "0"), ("decoy_prefix", "DECOY_"),
     ("equal_I_and_L", "1"), ("output_suffix", ""), ("mass_offsets", "")],
    [("minimum_peaks", "10"), ("minimum_intensity", "0"), ("remove_precursor_peak", "0"), ("remove_precursor_tolerance", "1.5"),
     ("clear_mz_range", "0.0 0.0")],
    [("add_Cterm_peptide", "0.0"), ("add_Nterm_peptide", "0.0"), ("add_Cterm_protein", "0.0"), ("add_Nterm_protein", "0.0")],
    [("fragment_bin_offset", "0")],
]
_COMET_ENZYMES = [("0.", "No_enzyme", "0", "-", "-"), ("1.", "Trypsin", "1", "KR", "P"), ("2.", "Trypsin/P", "1", "KR", "-"),
                  ("3.", "Lys_C", "1", "K", "P"), ("4.", "Lys_N", "0", "K", "-"), ("5.", "Arg_C", "1", "R", "P"),
                  ("6.", "Asp_N", "0", "D", "-"), ("7.", "CNBr", "1", "M", "-"), ("8.", "Glu_C", "1", "DE", "P"),
                  ("9.", "PepsinA", "1", "FL", "P"), ("10.", "Chymotrypsin", "1", "FWYL", "P")]


def _kv(key, value):
    return "%s = %s" % (key, value) if value != "" else "%s =" % key


def comet_static_modification_param(mod):
    """Modification::to_comet_static_modification_param (models/amino_acids/modification.rs:132-149)."""
    aa = mod.amino_acid.upper()
    if aa != "J":
        return "add_%s_%s = %s" % (aa, AMINO_ACID_NAMES[aa].lower(), rust_f64(mod.mono_mass_int / 1000000.0))
    return "add_J_user_amino_acid = %s" % rust_f64((J_MONO_INT + mod.mono_mass_int) / 1000000.0)


def comet_variable_modification_param(mod, number, max_variable_mods):
    """Modification::to_comet_variable_modification_param (modification.rs:151-171)."""
    if number > 9:
        raise ValueError("modification_number is not a number from 0 to 9")
    dist, term = {"A": (-1, 0), "C": (0, 3), "N": (0, 2)}[mod.position]
    return "variable_mod0%d = %s %s 0 %d %d %d 0" % (number, rust_f64(mod.mono_mass_int / 1000000.0), mod.amino_acid.upper(),
                                                     max_variable_mods, dist, term)


def comet_params(comet_revision, mods, fasta_path, n_targets_and_decoys, max_variable_mods, fragmentation_tolerance,
                 lower_ppm, upper_ppm):
    """comet_parameter::new (utility/comet_parameter.rs:96-124).  The reference iterates two HashMaps (random order);
    here fixed and variable modifications are emitted in ascending letter order."""
    out = [comet_revision, "", "# Comet MS/MS search engine parameters file.",
           "# Everything following the '#' symbol is treated as a comment.", ""]
    for group in _COMET_GROUPS_BEFORE:
        out.extend(_kv(k, v) for k, v in group)
        out.append("")
    text = "\n".join(out) + "\n"
    text += "peptide_mass_tolerance = %.4f\n" % float(max(lower_ppm, upper_ppm))
    text += "fragment_bin_tol = %s\n" % rust_f64(fragmentation_tolerance)
    text += "num_results = %d\n" % n_targets_and_decoys
    text += "num_output_lines = %d\n" % n_targets_and_decoys
    text += "database_name = %s\n" % fasta_path
    fix = sorted((m for m in mods if m.is_fix), key=lambda m: m.amino_acid)
    var = sorted((m for m in mods if not m.is_fix), key=lambda m: m.amino_acid)
    for m in fix:
        text += comet_static_modification_param(m) + "\n"
    if not any(m.amino_acid == "J" for m in fix):
        text += "add_J_user_amino_acid = 113.08406\n"
    for i, m in enumerate(var[:9]):
        text += comet_variable_modification_param(m, i + 1, max_variable_mods) + "\n"
    text += "\n[COMET_ENZYME_INFO]\n"
    for num, name, sense, cut, no_cut in _COMET_ENZYMES:
        text += "%-4s%-23s%-7s%-11s %s\n" % (num, name, sense, cut, no_cut)
    return text


def spectrum_fasta(target_rows, decoy_rows, mods):
    """The FASTA of one spectrum: targets then decoys, each deduplicated by sequence.
    target_rows: iterable of (sequence, peptide_id, var_mask); decoy_rows: iterable of (sequence, var_mask)."""
    seen, out = set(), []
    for seq, pid, mask in target_rows:
        if seq not in seen:
            seen.add(seq)
            out.append(fasta_entry(peptide_header(seq, pid, modification_summary(seq, mods, mask)), seq))
    seen_d = set()
    for seq, mask in decoy_rows:
        if seq not in seen_d:
            seen_d.add(seq)
            out.append(fasta_entry(decoy_header(seq, modification_summary(seq, mods, mask)), seq))
    return "".join(out), len(seen), len(seen_d)


def write_identification_outputs(directory, names, engine, spectra, params, mods, max_variable_mods, comet_revision):
    """identification_task's file outputs for every spectrum of a file (tasks/identification.rs:323-368): the candidate
    sets are exactly those that were scored.  names[i] = basename of spectrum i (the reference derives it from the
    one-spectrum mzML file name).  The file goes through the library in batches whose decoys fit one pass of the
    workspaces (2^25 decoy slots, 32k spectra); spectrum ids stay global, so the decoy RNG streams -- and with them the
    result -- do not depend on the batching.  Returns the PSM table."""
    import numpy as np
    from .api import SearchParams
    p = SearchParams(**{k: getattr(params, k) for k in ("lower_ppm", "upper_ppm", "fragment_tolerance", "n_decoys", "decoy_mode", "seed",
                                                         "top_k", "min_peaks", "max_fragment_charge", "abs_lower_uda", "abs_upper_uda")},
                     keep_decoys=True)
    table = engine.peptides()
    seqs = engine.sequences_of(table)
    os.makedirs(directory, exist_ok=True)
    n = len(spectra)
    batch = max(1, min(32768, (1 << 25) // max(1, p.n_decoys)))
    ids = spectra.spectrum_id if spectra.spectrum_id is not None else np.arange(n, dtype=np.uint32)
    all_psms, total = [], None
    for b0 in range(0, max(n, 1), batch):
        idx = np.arange(b0, min(n, b0 + batch))
        sub = spectra.subset(idx)
        sub.spectrum_id = np.ascontiguousarray(ids[idx], dtype=np.uint32)
        psms, stats = engine.identify(sub, p)
        all_psms.append(psms)
        total = stats if total is None else {k: total[k] + v for k, v in stats.items()}
        decoys = engine.last_decoys() if len(sub) else None
        pre = []
        for i in range(len(sub)):
            P, lo, hi = engine.precursor_window(float(sub.precursor_mz[i]), int(sub.charge[i]), p.lower_ppm, p.upper_ppm)
            if p.abs_lower_uda or p.abs_upper_uda:
                lo, hi = P - p.abs_lower_uda, P + p.abs_upper_uda
            pre.append((P, lo, hi, int(sub.charge[i]), int(sub.spectrum_id[i])))
        cand = engine.candidates(pre)
        raw, so = (decoys["seq"].tobytes(), decoys["seq_off"]) if decoys is not None else (b"", [0])
        for s in range(len(sub)):
            name = names[b0 + s]
            t = [(seqs[int(cand["peptide_id"][i]) - 1], int(cand["peptide_id"][i]), int(cand["var_mask"][i]))
                 for i in range(int(cand["off"][s]), int(cand["off"][s + 1]))]
            d = [(raw[int(so[i]):int(so[i + 1])].decode(), int(decoys["var_mask"][i]))
                 for i in range(int(decoys["off"][s]), int(decoys["off"][s + 1]))]
            text, nt, nd = spectrum_fasta(t, d, mods)
            fasta_path = os.path.join(directory, name + ".fasta")
            with open(fasta_path, "w") as fh:
                fh.write(text)
            if nd < p.n_decoys:                                   # GenerationResult::Timeout in the reference
                with open(os.path.join(directory, name + ".less_decoys"), "w") as fh:
                    fh.write("%d" % nd)
            with open(os.path.join(directory, name + ".comet.params"), "w") as fh:
                fh.write(comet_params(comet_revision, mods, fasta_path, nt + nd, max_variable_mods, p.fragment_tolerance, p.lower_ppm, p.upper_ppm))
    return np.concatenate(all_psms) if all_psms else np.zeros((0, p.top_k)), total
