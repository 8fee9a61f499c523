#!/usr/bin/env python
"""`max_decoy` command line: the reference's subcommands and flags (src/main.rs:213-496) on top of the C ABI.

The flags, their short forms and defaults are the reference's; what differs is the state between `digest` and
`identification`: the reference keeps it in PostgreSQL, here `identification` digests the FASTA given with
`--fasta` into the in-HBM index of the GPU it runs on (a human proteome takes < 1 s), and a whole mzML/MGF file is
identified in one call instead of one process per spectrum.  Decoys found per spectrum are written next to the
reference's own outputs (`<scan>.fasta`, `<scan>.comet.params`, `<scan>.less_decoys`) plus `psms.csv`.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _engine(args):
    import maxdecoy
    return maxdecoy.Engine(device=args.device)


def _read_proteins(path):
    from maxdecoy import synth
    with open(path) as fh:
        return synth.read_fasta(fh.read())


def _read_spectra(path):
    from maxdecoy import mzml, synth
    if path.lower().endswith(".mgf"):
        with open(path) as fh:
            sp = synth.read_mgf(fh.read())
        return sp, [("scan=%d" % (i + 1), str(i + 1)) for i in range(len(sp))]
    return mzml.read_ms_two_spectra(path)


def cmd_digest(args):
    """`digest` (src/main.rs:214-272, tasks/digestion.rs:43-136): peptides / peptides_proteins / proteins as COPY-able CSV."""
    from maxdecoy import pgexport
    if args.max_len > 60 or args.missed > 60:
        sys.exit("maximum peptide length and missed cleavages must be <= 60 (tasks/digestion.rs:74,101)")
    headers, seqs = _read_proteins(args.input_file)
    eng = _engine(args)
    n = eng.digest(seqs, args.missed, args.min_len, args.max_len)
    table = eng.peptides()
    os.makedirs(args.out, exist_ok=True)
    for name, text in (("proteins.csv", pgexport.proteins_csv(headers, seqs)), ("peptides.csv", pgexport.peptides_csv(table)),
                       ("peptides_proteins.csv", pgexport.peptides_proteins_csv(table))):
        with open(os.path.join(args.out, name), "w") as fh:
            fh.write(text)
    print("%d proteins, %d unique peptides -> %s" % (len(seqs), n, args.out))


def _read_stored_decoys(path):
    """The `decoys` table as CSV (`pgexport.decoys_csv`: id, aa_sequence, ...) or one sequence per line: the decoys the
    reference would find in its database and reuse (tasks/identification.rs:259-283)."""
    out = []
    with open(path) as fh:
        for line in fh:
            f = [x.strip().strip('"') for x in line.strip().split(",")]
            q = f[1] if len(f) > 1 else f[0]
            if q and q.isalpha() and q.isupper():
                out.append(q)
    return out


def cmd_identification(args):
    """`identification` (src/main.rs:397-485, tasks/identification.rs:160-370) for every MS2 spectrum of the file."""
    import maxdecoy
    from maxdecoy import outputs, pgexport
    mods = maxdecoy.Modification.create_from_csv_file(args.modification_file)
    _, seqs = _read_proteins(args.fasta)
    spectra, ids = _read_spectra(args.spectrum_file)
    eng = _engine(args)
    eng.digest(seqs, args.missed, args.min_len, args.max_len)
    eng.set_modifications(mods, args.nvar)
    eng.set_variable_mode(maxdecoy.VARMOD_EXPANDED if args.variable_mode == "expanded" else maxdecoy.VARMOD_REFERENCE)
    if args.stored_decoys:
        eng.set_decoy_store(_read_stored_decoys(args.stored_decoys))
    eng.index_build()
    prm = maxdecoy.SearchParams(args.lower, args.upper, fragment_tolerance=args.fragmentation_tolerance, n_decoys=args.decoys,
                                decoy_mode=args.decoy_mode, seed=args.seed, top_k=args.top_k)
    names = [(sid[1] or sid[0]).replace("/", "_").replace(" ", "_") for sid in ids]
    psms, stats = outputs.write_identification_outputs(args.out, names, eng, spectra, prm, mods, args.nvar, args.comet_revision)
    table_seqs = eng.sequences_of(eng.peptides())
    dec = eng.last_decoys()
    raw, so = dec["seq"].tobytes(), dec["seq_off"]

    def seq_of(s, row):
        if row["is_decoy"]:
            i = int(dec["off"][s]) + int(row["candidate"])
            return raw[int(so[i]):int(so[i + 1])].decode()
        return table_seqs[int(row["candidate"]) - 1]
    pm = [eng.precursor_window(float(spectra.precursor_mz[i]), int(spectra.charge[i]), args.lower, args.upper)[0] for i in range(len(spectra))]
    with open(os.path.join(args.out, "decoys.csv"), "w") as fh:       # rows of table `decoys`: what a later run can reuse (--stored-decoys)
        fh.write(pgexport.decoys_csv(dec))
    with open(os.path.join(args.out, "psms.csv"), "w") as fh:
        fh.write(pgexport.psms_csv(psms, ids, seq_of, lambda s, row: outputs.modification_summary(seq_of(s, row), mods, int(row["var_mask"])), pm))
    print("%d spectra, %d targets, %d decoys scored, %d spectra with fewer decoys than requested -> %s"
          % (stats["n_spectra"], stats["n_targets"], stats["n_decoys"], stats["n_less_decoys"], args.out))


def cmd_decoy_generation(args):
    """`decoy-generation` (src/main.rs:274-340): decoys for one precursor mass (Da) and a ppm window.
    (The reference passes the ppm integers as absolute limits, in swapped order -- main.rs:126-133; fixed here.)"""
    import maxdecoy
    from maxdecoy import mass
    mods = maxdecoy.Modification.create_from_csv_file(args.modification_file)
    _, seqs = _read_proteins(args.fasta) if args.fasta else ([], [])
    eng = _engine(args)
    eng.digest(seqs, 2, 5, 50)
    eng.set_modifications(mods, args.nvar)
    eng.index_build()
    P = mass.convert_mass_to_int(args.precursor_mass)
    lo, hi = P - P * args.lower // 1000000, P + P * args.upper // 1000000
    d = eng.generate_decoys([(P, lo, hi, 2, 0)], args.decoys, args.decoy_mode, args.seed)
    raw, so = d["seq"].tobytes(), d["seq_off"]
    for i in range(len(so) - 1):
        print(raw[int(so[i]):int(so[i + 1])].decode())


def cmd_sequence_mass(args):
    """`sequence-mass` (tasks/sequence_mass.rs:24-27)."""
    import maxdecoy
    lib = maxdecoy.load()
    b = args.sequence.encode()
    print(lib.md_sequence_weight(b, len(b)) / 1000000.0)


def cmd_substitution(args):
    """`amino-acid-substitution` (src/main.rs:140-180): mass difference of one substitution incl. fixed modifications."""
    import maxdecoy
    from maxdecoy._abi import ALPHABET
    mods = maxdecoy.Modification.create_from_csv_file(args.modification_file)
    eng = _engine(args)
    eng.set_modifications(mods, 0)
    m = eng.substitution_map()
    print(int(m[ALPHABET.index(args.source.upper()), ALPHABET.index(args.destination.upper())]) / 1000000.0)


def cmd_spectrum_splitup(args):
    """`spectrum-splitup` (src/main.rs:183-206): one mzML per MS2 spectrum (kept for users who still run Comet)."""
    from maxdecoy import mzml
    with open(args.mz_ml_file) as fh:
        names = mzml.spectrum_splitup(fh.read(), args.destination_folder, args.file_suffix)
    print("found and write %d MS2-spectra -> %s" % (len(names), args.destination_folder))


def main(argv=None):
    ap = argparse.ArgumentParser(prog="max_decoy", add_help=False)
    ap.add_argument("--help", action="help")
    ap.add_argument("--device", type=int, default=0)
    sub = ap.add_subparsers(dest="cmd", required=True)

    p = sub.add_parser("digest", add_help=False)
    p.add_argument("--help", action="help")
    p.add_argument("-i", "--input-file", required=True)
    p.add_argument("-f", "--format", default="fasta", choices=["fasta"])
    p.add_argument("-t", "--thread-count", type=int, default=2, help="accepted for compatibility; the GPU parallelises")
    p.add_argument("-c", "--number-of-missed-cleavages", dest="missed", type=int, default=2)
    p.add_argument("-l", "--minimum-peptide_length", dest="min_len", type=int, default=5)
    p.add_argument("-h", "--maximum-peptide_length", dest="max_len", type=int, default=50)
    p.add_argument("-e", "--enzym-name", default="Trypsin")
    p.add_argument("-o", "--out", default="digest_out")
    p.set_defaults(fn=cmd_digest)

    def common_decoy(p):
        p.add_argument("--help", action="help")
        p.add_argument("-m", "--modification-file", required=True)
        p.add_argument("-d", "--number-of-decoys", dest="decoys", type=int, default=1000)
        p.add_argument("-l", "--lower-mass-tolerance", dest="lower", type=int, default=5)
        p.add_argument("-u", "--upper-mass-tolerance", dest="upper", type=int, default=5)
        p.add_argument("-t", "--thread-count", type=int, default=2, help="accepted for compatibility")
        p.add_argument("--max-time-for-decoy-generation", type=int, default=60, help="accepted for compatibility: attempts are bounded, not time")
        p.add_argument("--decoy-mode", type=int, default=0, help="0 reference-random, 1 exhaustive, 2 permuted targets")
        p.add_argument("--seed", type=int, default=0)
        p.add_argument("--fasta", help="protein FASTA (replaces the PostgreSQL peptide table)")

    p = sub.add_parser("decoy-generation", add_help=False)
    common_decoy(p)
    p.add_argument("-n", "--max-modification-per-decoy", dest="nvar", type=int, default=0)
    p.add_argument("-p", "--precursor-mass", type=float, required=True)
    p.set_defaults(fn=cmd_decoy_generation)

    p = sub.add_parser("identification", add_help=False)
    common_decoy(p)
    p.add_argument("-s", "--spectrum-file", required=True, help="mzML (reference format) or MGF")
    p.add_argument("-n", "--max-number-of-variable-modification-per-peptide", dest="nvar", type=int, default=0)
    p.add_argument("--fragmentation-tolerance", type=float, default=0.02)
    p.add_argument("-r", "--comet-revision", default="# comet_version 2019.01 rev. 4")
    p.add_argument("-c", "--number-of-missed-cleavages", dest="missed", type=int, default=2)
    p.add_argument("--minimum-peptide_length", dest="min_len", type=int, default=5)
    p.add_argument("--maximum-peptide_length", dest="max_len", type=int, default=50)
    p.add_argument("--top-k", type=int, default=5)
    p.add_argument("--variable-mode", choices=["reference", "expanded"], default="reference",
                   help="reference: all-or-first-hit placement like the reference; expanded: every placement of <= n variable modifications")
    p.add_argument("--stored-decoys", default="", help="CSV of the `decoys` table (or one sequence per line): reused before new decoys are generated")
    p.add_argument("-o", "--out", default="identification_out")
    p.set_defaults(fn=cmd_identification)

    p = sub.add_parser("spectrum-splitup", add_help=False)
    p.add_argument("--help", action="help")
    p.add_argument("-m", "--mz-ml-file", required=True)
    p.add_argument("-d", "--destination-folder", required=True)
    p.add_argument("-s", "--file-suffix", default="")
    p.set_defaults(fn=cmd_spectrum_splitup)

    p = sub.add_parser("amino-acid-substitution", add_help=False)
    p.add_argument("--help", action="help")
    p.add_argument("-m", "--modification_file", dest="modification_file", required=True)
    p.add_argument("-s", "--source-amino-acid", dest="source", required=True)
    p.add_argument("-d", "--destination-amino-acid", dest="destination", required=True)
    p.set_defaults(fn=cmd_substitution)

    p = sub.add_parser("sequence-mass", add_help=False)
    p.add_argument("--help", action="help")
    p.add_argument("-s", "--sequence", required=True)
    p.set_defaults(fn=cmd_sequence_mass)

    args = ap.parse_args(argv)
    args.fn(args)


if __name__ == "__main__":
    main()
