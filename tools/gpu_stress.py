"""Randomized parity stress (GPU vs oracle; test infrastructure, run by tests/test_gpu_stress.py): decoys, candidates and identify over random seeds, modification
sets, decoy counts, windows and batch sizes.  python tools/gpu_stress.py [rounds [seed]]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "max-decoy_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import maxdecoy
from maxdecoy import SearchParams, synth, Modification
from oracle_lib import oracle_engine

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
gpu, cpu = maxdecoy.Engine(), oracle_engine(8)
extra = [Modification("x:21", "Phospho", "A", False, "S", 79.966331), Modification("x:7", "Deamid", "A", False, "N", 0.984016),
         Modification("x:9", "FixK", "A", True, "K", 8.014199)]
# terminal (N / C) modifications, drawn from their own stream so that the other configurations stay what they were
terminal = [Modification("t:1", "Acetyl", "N", True, "M", 42.010565), Modification("t:2", "PyroGlu", "N", False, "Q", -17.026549),
            Modification("t:3", "LabelK", "C", True, "K", 8.014199), Modification("t:4", "MethylR", "C", False, "R", 14.01565),
            Modification("t:5", "CarbS", "N", False, "S", 43.005814)]
trng = np.random.default_rng((int(sys.argv[2]) if len(sys.argv) > 2 else 2026) + 1)
bad = 0
for r in range(rounds):
    n_prot = int(rng.integers(50, 500)); n_spec = int(rng.integers(1, 70)); mc = int(rng.integers(0, 3))
    mods = [synth.CAM] + ([synth.OXM] if rng.random() < 0.6 else []) + [m for m in extra if rng.random() < 0.3]
    nvar = int(rng.integers(0, 4)); nd = int(rng.choice([0, 1, 7, 64, 300])); mode = int(rng.choice([0, 0, 0, 2]))
    topk = int(rng.choice([1, 5, 8, 20])); ppm = int(rng.choice([5, 10, 50])); absw = int(rng.choice([0, 0, 0, 3_000_000, 40_000_000]))
    expanded = rng.random() < 0.3
    term = [m for m in terminal if trng.random() < 0.35] if trng.random() < 0.35 else []
    mods = mods + term    # (decoy modes here: 0 reference-random, 2 permuted targets; exhaustive decoys are not defined with terminal modifications)
    prots = synth.synthetic_proteins(n_prot, seed=int(rng.integers(1 << 30)))
    sp, _ = synth.synthetic_spectra(prots, n_spec, mc, mods=tuple(m for m in mods if m.amino_acid in "CM"), seed=int(rng.integers(1 << 30)))
    for e in (gpu, cpu):
        e.digest(prots, mc, 5, 50); e.set_modifications(mods, nvar)
        e.set_variable_mode(maxdecoy.VARMOD_EXPANDED if expanded else maxdecoy.VARMOD_REFERENCE); e.index_build()
    store = []
    if rng.random() < 0.4 and mode == 0 and nd:
        pre = [tuple(cpu.precursor_window(float(sp.precursor_mz[i]), int(sp.charge[i]), ppm, ppm)) + (int(sp.charge[i]), i) for i in range(len(sp))]
        d0 = cpu.generate_decoys(pre, 20, 0, seed=r)
        raw, so = d0["seq"].tobytes(), d0["seq_off"]
        store = [raw[int(so[i]):int(so[i + 1])].decode() for i in range(0, len(so) - 1, 2)]
    for e in (gpu, cpu):
        e.set_decoy_store(store)
    prm = SearchParams(ppm, ppm, n_decoys=nd, decoy_mode=mode, seed=int(rng.integers(1 << 40)), top_k=topk, abs_lower_uda=absw, abs_upper_uda=absw)
    if absw > 3_000_000 and rng.random() < 0.5:
        os.environ["MD_SCORE_SPLIT_MIN"] = "1"
    else:
        os.environ.pop("MD_SCORE_SPLIT_MIN", None)
    pg, sg, scg, og = gpu.identify(sp, prm, want_all_scores=True)
    pc, sc, scc, oc = cpu.identify(sp, prm, want_all_scores=True)
    ok = np.array_equal(og, oc) and np.array_equal(scg, scc) and all(np.array_equal(pg[k], pc[k]) for k in pg.dtype.names if k != "_pad")
    if not ok:
        print("   offsets equal:", np.array_equal(og, oc), "scores equal:", np.array_equal(scg, scc), "split:", os.environ.get("MD_SCORE_SPLIT_MIN"),
              [k for k in pg.dtype.names if k != "_pad" and not np.array_equal(pg[k], pc[k])])
        if np.array_equal(og, oc) and not np.array_equal(scg, scc):
            w = np.nonzero(scg != scc)[0]
            spec = np.searchsorted(og, w, side="right") - 1
            print("   first differing candidates:", w[:8], "spectra:", spec[:8], "of n_targets", pg["n_targets"][spec[:8], 0], "gpu", scg[w[:4]], "cpu", scc[w[:4]])
    print("round %2d prot=%d spec=%d mc=%d mods=%d (terminal %d) nvar=%d nd=%d mode=%d k=%d ppm=%d abs=%d exp=%d store=%d pairs=%d -> %s"
          % (r, n_prot, n_spec, mc, len(mods), len(term), nvar, nd, mode, topk, ppm, absw, expanded, len(store), len(scg), "ok" if ok else "MISMATCH"), flush=True)
    bad += not ok
print("mismatches:", bad)
sys.exit(1 if bad else 0)
