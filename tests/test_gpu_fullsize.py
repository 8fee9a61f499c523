"""BASELINE.json's full single-GPU configuration (C2: 20k proteins, MC=2, 10k spectra, 1000 decoys/spectrum) through
size-independent properties -- the oracle cannot score 10.8 M pairs in seconds, so at this size the checks are:
digest equal to the oracle's (the oracle digests 20k proteins in seconds), decoy properties on every decoy, PSM order,
a sample of spectra re-scored by the oracle bit for bit, and invariance under batching (what multi-GPU sharding relies on)."""
import numpy as np
import pytest

import maxdecoy
from maxdecoy import SearchParams, parallel, synth
from oracle_lib import oracle_engine

pytestmark = pytest.mark.gpu

N_PROT, N_SPEC, N_DECOY = 20000, 10000, 1000


@pytest.fixture(scope="module")
def world():
    prots = synth.synthetic_proteins(N_PROT)
    sp, truth = synth.synthetic_spectra(prots, N_SPEC, 2, seed=7)
    gpu = maxdecoy.Engine()
    cpu = oracle_engine(16)
    for e in (gpu, cpu):
        e.digest(prots, 2, 5, 50)
        e.set_modifications([synth.CAM], 0)
        e.index_build()
    yield gpu, cpu, sp, truth
    gpu.close()
    cpu.close()


def test_digest_and_index_equal_the_oracle(world):
    gpu, cpu, _, _ = world
    tg, tc = gpu.peptides(), cpu.peptides()
    assert len(tg["weight"]) > 2_000_000
    for k in tg:
        assert np.array_equal(tg[k], tc[k]), k
    n = gpu.index_stats()["n_peptides"]
    pg, kg = gpu.index_export(0, n)
    pc, kc = cpu.index_export(0, n)
    assert np.array_equal(pg, pc) and np.array_equal(kg, kc)


def test_identify_properties_at_full_size(world):
    gpu, cpu, sp, truth = world
    prm = SearchParams(10, 10, n_decoys=N_DECOY, seed=20260101, top_k=5, keep_decoys=True)
    psms, st = gpu.identify(sp, prm)
    assert st["n_spectra"] == N_SPEC and st["n_pairs"] == st["n_targets"] + st["n_decoys"] > 10_000_000
    # PSM rows: ranks 1..k contiguous, raw score descending, ties by candidate order (targets before decoys)
    rk = psms["rank"].astype(np.int64)
    filled = rk > 0
    assert np.all(filled[:, :-1] >= filled[:, 1:]) and np.all(rk[filled] == (np.nonzero(filled)[1] + 1))
    sc = psms["raw_score"]
    assert np.all((sc[:, :-1] >= sc[:, 1:]) | ~filled[:, 1:])
    assert np.array_equal(psms["spectrum_id"][:, 0], np.arange(N_SPEC, dtype=np.uint32))
    # most database spectra are won by their generating peptide
    seqs = gpu.sequences_of(gpu.peptides())
    top = psms[:, 0]
    won = sum(1 for i in range(0, N_SPEC, 10) if top["rank"][i] and not top["is_decoy"][i] and seqs[int(top["candidate"][i]) - 1] == truth[i][0])
    assert won > 0.75 * (N_SPEC // 10)
    # every decoy: alphabet, modified weight inside the window, not a peptide, unique within its spectrum, <= 60 residues
    d = gpu.last_decoys()
    assert len(d["attempt"]) == st["n_decoys"]
    raw, so, off = d["seq"].tobytes(), d["seq_off"].astype(np.int64), d["off"].astype(np.int64)
    lens = np.diff(so)
    assert lens.min() >= 1 and lens.max() <= 60
    assert set(np.unique(d["seq"]).tolist()) <= set(maxdecoy.ALPHABET.encode())
    targets = set(seqs)
    P = np.zeros(N_SPEC, dtype=np.int64); lo = P.copy(); hi = P.copy()
    for i in range(N_SPEC):
        P[i], lo[i], hi[i] = gpu.precursor_window(float(sp.precursor_mz[i]), int(sp.charge[i]), 10, 10)
    spec_of = np.repeat(np.arange(N_SPEC), np.diff(off))
    assert np.all(d["mod_weight"] >= lo[spec_of]) and np.all(d["mod_weight"] <= hi[spec_of])
    cam = np.frombuffer(raw, dtype=np.uint8) == ord("C")
    ncys = np.add.reduceat(cam.astype(np.int64), so[:-1]) if len(so) > 1 else np.zeros(0, dtype=np.int64)
    assert np.array_equal(d["mod_weight"], d["weight"] + ncys * 57021464)          # fixed CAM on every C, nothing else
    for s in range(0, N_SPEC, 25):                                                  # string-level checks on a sample
        mine = [raw[so[i]:so[i + 1]].decode() for i in range(off[s], off[s + 1])]
        assert len(set(mine)) == len(mine) and not (set(mine) & targets)
        assert np.all(np.diff(d["attempt"][off[s]:off[s + 1]].astype(np.int64)) > 0)
    # a sample of spectra re-identified by the oracle: bit for bit
    idx = np.arange(0, N_SPEC, 40)            # 250 spectra
    sub = sp.subset(idx)
    pc, _ = cpu.identify(sub, prm)
    for f in ("spectrum_id", "rank", "is_decoy", "candidate", "var_mask", "mod_weight", "raw_score", "n_targets", "n_decoys"):
        assert np.array_equal(psms[f][idx], pc[f]), f
    # batching / sharding invariance: two mass-interleaved halves identified separately give the same rows
    parts = parallel.partition_spectra(sp.precursor_mz, sp.charge, 2)
    merged = np.zeros_like(psms)
    for part in parts:
        sub = sp.subset(part)
        pp, _ = gpu.identify(sub, prm)
        merged[part] = pp
    assert merged.tobytes() == psms.tobytes()


def test_c5_open_search_at_full_size(world):
    """BASELINE configs[4] on the real index: +-500 Da windows over the 20k-protein index (~630k targets per spectrum, no
    decoys); fewer spectra than SMs, so every spectrum is split into parts over the SMs (no environment override).  Every
    PSM row and every raw score equal the oracle's."""
    gpu, cpu, sp, _ = world
    idx = np.array([3, 1500, 4200, 6100, 7777, 9990])
    sub = sp.subset(idx)
    prm = SearchParams(10, 10, n_decoys=0, seed=1, top_k=5, abs_lower_uda=500_000_000, abs_upper_uda=500_000_000)
    pg, sg, scg, og = gpu.identify(sub, prm, want_all_scores=True)
    pc, sc, scc, oc = cpu.identify(sub, prm, want_all_scores=True)
    assert sg["n_targets"] == sc["n_targets"] > 6 * 300_000
    assert np.array_equal(og, oc) and np.array_equal(scg, scc)
    for f in ("spectrum_id", "rank", "is_decoy", "candidate", "var_mask", "mod_weight", "raw_score", "n_targets", "n_decoys"):
        assert np.array_equal(pg[f], pc[f]), f
    # PSM rows only (the path bench.py times): the same rows without the all-scores side output
    pg2, _ = gpu.identify(sub, prm)
    assert pg2.tobytes() == pg.tobytes()


def test_c3_variable_modifications_at_full_size(world):
    """BASELINE configs[2]: the 20k-protein index with fixed CAM-C + variable Met-oxidation (<= 3 per peptide), 10k spectra,
    1000 mass-matched decoys each: properties of EVERY decoy, and 100 spectra re-identified by the oracle bit for bit."""
    gpu, cpu, _, _ = world
    prots = synth.synthetic_proteins(N_PROT)
    sp, _ = synth.synthetic_spectra(prots, N_SPEC, 2, mods=(synth.CAM, synth.OXM), seed=7)
    for e in (gpu, cpu):
        e.set_modifications([synth.CAM, synth.OXM], 3)
        e.index_build()
    try:
        prm = SearchParams(10, 10, n_decoys=N_DECOY, seed=20260101, top_k=5, keep_decoys=True)
        psms, st = gpu.identify(sp, prm)
        assert st["n_spectra"] == N_SPEC and st["n_decoys"] > 9_000_000
        d = gpu.last_decoys()
        raw, so, off = np.frombuffer(d["seq"].tobytes(), dtype=np.uint8), d["seq_off"].astype(np.int64), d["off"].astype(np.int64)
        lens = np.diff(so)
        assert lens.min() >= 1 and lens.max() <= 60
        assert set(np.unique(raw).tolist()) <= set(maxdecoy.ALPHABET.encode())
        P = np.zeros(N_SPEC, dtype=np.int64); lo = P.copy(); hi = P.copy()
        for i in range(N_SPEC):
            P[i], lo[i], hi[i] = gpu.precursor_window(float(sp.precursor_mz[i]), int(sp.charge[i]), 10, 10)
        spec_of = np.repeat(np.arange(N_SPEC), np.diff(off))
        assert np.all(d["mod_weight"] >= lo[spec_of]) and np.all(d["mod_weight"] <= hi[spec_of])
        ncys = np.add.reduceat((raw == ord("C")).astype(np.int64), so[:-1])
        vm = d["var_mask"]
        nvar = np.array([bin(int(m)).count("1") for m in vm[::97]])           # popcounts on a sample (python ints)
        assert nvar.max() <= 3
        pop = np.zeros(len(vm), dtype=np.int64)
        m = vm.copy()
        while m.any():
            pop += (m & np.uint64(1)).astype(np.int64); m >>= np.uint64(1)
        assert pop.max() <= 3
        assert np.array_equal(d["mod_weight"], d["weight"] + ncys * 57021464 + pop * 15994915)
        # a variable modification sits on an M only
        for k in range(0, len(vm), 5003):
            mk = int(vm[k]); s = raw[so[k]:so[k + 1]]
            assert all(s[i] == ord("M") for i in range(len(s)) if (mk >> i) & 1)
        for s in range(0, N_SPEC, 50):
            mine = [raw[so[i]:so[i + 1]].tobytes() for i in range(off[s], off[s + 1])]
            assert len(set(mine)) == len(mine)
        idx = np.arange(0, N_SPEC, 100)           # 100 spectra
        pc, _ = cpu.identify(sp.subset(idx), prm)
        for f in ("spectrum_id", "rank", "is_decoy", "candidate", "var_mask", "mod_weight", "raw_score", "n_targets", "n_decoys"):
            assert np.array_equal(psms[f][idx], pc[f]), f
    finally:
        for e in (gpu, cpu):
            e.set_modifications([synth.CAM], 0)
            e.index_build()


def test_c1_every_psm_and_score_equal_the_oracle():
    """BASELINE configs[0] in full (the CPU-runnable case): 2k proteins, MC=1, fixed CAM-C, 1k spectra, 10 ppm, 1000
    decoys per spectrum -- every PSM row and the raw score of every one of the ~1.05 M candidates equal the oracle's."""
    prots = synth.synthetic_proteins(2000)
    sp, _ = synth.synthetic_spectra(prots, 1000, 1, seed=7)
    prm = SearchParams(10, 10, n_decoys=1000, seed=20260101, top_k=5)
    out = []
    for e in (maxdecoy.Engine(), oracle_engine(16)):
        e.digest(prots, 1, 5, 50)
        e.set_modifications([synth.CAM], 0)
        e.index_build()
        out.append(e.identify(sp, prm, want_all_scores=True))
        e.close()
    (pg, sg, scg, og), (pc, sc, scc, oc) = out
    assert sg["n_targets"] == sc["n_targets"] and sg["n_decoys"] == sc["n_decoys"] > 900_000
    assert np.array_equal(og, oc) and np.array_equal(scg, scc)
    assert pg.tobytes() == pc.tobytes()
