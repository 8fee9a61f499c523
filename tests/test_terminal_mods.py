"""Terminal (N / C position) modifications, SURVEY 8(f) row 4.

Definition: the slot model the reference's add_modification_at / set_variable_modification_at
spell out (modified_peptide.rs:421-447, :339-367) -- a letter's terminal modification sits on the first / last residue
only, and only when that residue is the letter.  The SQL fan-out stays position-blind as in the reference
(identification.rs:190-222 shifts the window by count x delta for every modified letter), the substitution map of the
decoy repair loop too (decoy_generator.rs:301-324).

CPU part: the oracle (functional restatement) against the literal slot-by-slot Python restatement (pyref.SlotPeptide).
GPU part: CUDA against the oracle, bit for bit.
"""
import numpy as np
import pytest

import maxdecoy
from maxdecoy import Modification, SearchParams, synth
from oracle_lib import oracle_engine
import pyref
import workloads as wl

K_CTERM_FIX = Modification("x:259", "Label:13C(6)15N(2)", "C", True, "K", 8.014199)
Q_NTERM_VAR = Modification("unimod:28", "Gln->pyro-Glu", "N", False, "Q", -17.026549)
M_NTERM_FIX = Modification("unimod:1", "Acetyl", "N", True, "M", 42.010565)
R_CTERM_VAR = Modification("unimod:34", "Methyl", "C", False, "R", 14.01565)
K_ANY_FIX = Modification("x:1", "Dimethyl", "A", True, "K", 28.0313)
K_CTERM_VAR = Modification("x:2", "Label", "C", False, "K", 8.014199)
S_NTERM_FIX = Modification("x:3", "Acetyl", "N", True, "S", 42.010565)
S_NTERM_VAR = Modification("x:4", "Blocked", "N", False, "S", 14.01565)   # same slot as the fixed one: never applies

MOD_SETS = [
    ((K_CTERM_FIX,), 0),
    ((synth.CAM, Q_NTERM_VAR), 2),
    ((M_NTERM_FIX, synth.OXM), 2),                    # N-terminal fixed and side-chain variable on the same letter
    ((K_ANY_FIX, K_CTERM_VAR, R_CTERM_VAR), 2),       # side-chain fixed and C-terminal variable on the same letter
    ((S_NTERM_FIX, S_NTERM_VAR, K_CTERM_FIX, synth.OXM), 3),
]
IDS = ["fixC", "varN", "fixN+varA", "fixA+varC", "mixed"]


@pytest.fixture(scope="module")
def cpu():
    e = oracle_engine(8)
    yield e
    e.close()


def _setup(e, n_prot, mods, nvar):
    e.digest(list(wl.proteins(n_prot)), 2, 5, 50)
    e.set_modifications(list(mods), nvar)
    e.index_build()


def _table(e):
    t = e.peptides()
    return [(s, int(w), list(c)) for s, w, c in zip(e.sequences_of(t), t["weight"], t["counts"])]


def _precursors(pm, peptides, seed, n=24, narrow=12):
    """Precursors that the slot-model weight of a sampled peptide (with a random legal variable placement) hits: `narrow`
    of them with a 10 ppm window, the rest 20 Da wide (several placements / many neighbours inside)."""
    rng = np.random.default_rng(seed)
    pre = []
    while len(pre) < n:
        seq = peptides[int(rng.integers(len(peptides)))][0]
        mp = pyref.SlotPeptide(pm, seq)
        elig = [i for i, c in enumerate(seq) if c in pm.var]
        rng.shuffle(elig)
        k = 0
        for i in elig:
            if k < pm.nvar and rng.random() < 0.7 and mp.set_variable(i):
                k += 1
        P = mp.w + int(rng.integers(-3, 4))
        tol = P // 100_000 if len(pre) < narrow else 20_000_000
        pre.append((P, P - tol, P + tol, int(rng.integers(2, 4)), len(pre)))
    return pre


def _spectra_for(e, peptides, mods, nvar, n, seed):
    """Synthetic MS2 spectra of sampled peptides under the slot model (b / y ladders with the terminal masses on the end residues)."""
    pm = pyref.Mods(mods, nvar)
    rng = np.random.default_rng(seed)
    pmz, charge, off, mzs, ints = [], [], [0], [], []
    while len(pmz) < n:
        seq = peptides[int(rng.integers(len(peptides)))][0]
        if not (7 <= len(seq) <= 40):
            continue
        mp = pyref.SlotPeptide(pm, seq)
        for i in [i for i, c in enumerate(seq) if c in pm.var][:nvar]:
            mp.set_variable(i)
        mask = mp.var_mask()
        last = len(seq) - 1
        res = []
        for i, c in enumerate(seq):
            v = pyref.residue_mass(c)
            fp = pm.fix_pos.get(c, "A")
            if c in pm.fix and (fp == "A" or (fp == "N" and i == 0) or (fp == "C" and i == last)):
                v += pm.fix[c]
            if (mask >> i) & 1:
                v += pm.var[c]
            res.append(v / 1e6)
        z = int(rng.choice([2, 3]))
        M = mp.w / 1e6
        pmz.append((M + z * synth.PROTON) / z)
        charge.append(z)
        pre = np.cumsum(res)
        peaks = []
        for k in range(1, len(seq)):
            b, y = pre[k - 1], pre[-1] - pre[k - 1] + synth.H2O
            for m in (b, y):
                if rng.random() < 0.8:
                    peaks.append((m + synth.PROTON + rng.normal(0.0, 0.004), float(rng.lognormal(5.0, 1.0))))
        for _ in range(60):
            peaks.append((float(rng.uniform(100.0, 1800.0)), float(rng.lognormal(4.0, 1.0))))
        peaks.sort()
        mzs.extend(p[0] for p in peaks)
        ints.extend(p[1] for p in peaks)
        off.append(len(mzs))
    return synth.Spectra(np.array(pmz), np.array(charge, dtype=np.uint8), np.array(off, dtype=np.uint64), np.array(mzs, dtype=np.float64),
                         np.array(ints, dtype=np.float32))


# ------------------------------------------------------------------------------------------------ CPU: oracle vs literal restatement
@pytest.mark.parametrize("mods,nvar", MOD_SETS, ids=IDS)
def test_terminal_candidates_equal_literal_slots(cpu, mods, nvar):
    pm = pyref.Mods(mods, nvar)
    heavy = len(pm.letters) > 2                      # the literal fan-out is K^letters table scans per precursor
    _setup(cpu, 12 if heavy else 60, mods, nvar)
    peptides = _table(cpu)
    pre = _precursors(pm, peptides, seed=11, n=8 if heavy else 24, narrow=4 if heavy else 12)
    got = cpu.candidates(pre)
    n_found = n_var = 0
    for s, (P, lo, hi, z, sid) in enumerate(pre):
        want = pyref.candidates_sql(pm, peptides, P, lo, hi)
        a, b = int(got["off"][s]), int(got["off"][s + 1])
        have = {int(got["peptide_id"][i]) - 1: (int(got["mod_weight"][i]), int(got["var_mask"][i])) for i in range(a, b)}
        assert have == want, s
        n_found += len(want)
        n_var += sum(1 for w, m in want.values() if m)
    assert n_found > 0
    if nvar and mods != MOD_SETS[4][0]:
        assert n_var > 0


def test_terminal_random_modification_sets(cpu):
    """Random modification sets (letters, positions A / N / C, fixed / variable, signs) on two letters: the oracle's
    candidates against the literal fan-out + slot-by-slot filter, in reference mode; and the expanded mode against the
    brute force over all placements."""
    rng = np.random.default_rng(99)
    letters = "KRMQSCNDE"
    n_found = n_exp = 0
    for trial in range(24):
        two = rng.choice(len(letters), size=2, replace=False)
        mods = []
        for li in two:
            for is_fix in (True, False):
                if rng.random() < 0.65:
                    delta = float(rng.choice([42.010565, 15.994915, 8.014199, -17.026549, 14.01565, 79.966331]))
                    mods.append(Modification("r:%d%d" % (li, is_fix), "m%d%d" % (li, is_fix), str(rng.choice(["A", "N", "C"])), is_fix, letters[li], delta))
        if not mods:
            continue
        nvar = int(rng.integers(0, 4))
        _setup(cpu, 25, mods, nvar)
        peptides = _table(cpu)
        pm = pyref.Mods(mods, nvar)
        pre = _precursors(pm, peptides, seed=100 + trial, n=6, narrow=3)
        got = cpu.candidates(pre)
        for s, (P, lo, hi, z, sid) in enumerate(pre):
            want = pyref.candidates_sql(pm, peptides, P, lo, hi)
            a, b = int(got["off"][s]), int(got["off"][s + 1])
            have = {int(got["peptide_id"][i]) - 1: (int(got["mod_weight"][i]), int(got["var_mask"][i])) for i in range(a, b)}
            assert have == want, (trial, s, [(m.amino_acid, m.position, m.is_fix) for m in mods])
            n_found += len(want)
        if nvar and pm.var:
            cpu.set_variable_mode(maxdecoy.VARMOD_EXPANDED)
            try:
                cpu.index_build()
                exp = cpu.candidates(pre[:3])
                seqs = [p[0] for p in peptides]
                for s, (P, lo, hi, z, sid) in enumerate(pre[:3]):
                    a, b = int(exp["off"][s]), int(exp["off"][s + 1])
                    got_e = {(int(exp["peptide_id"][i]) - 1, int(exp["var_mask"][i]), int(exp["mod_weight"][i])) for i in range(a, b)}
                    assert got_e == _expanded_bruteforce(pm, seqs, lo, hi), (trial, s)
                    n_exp += len(got_e)
            finally:
                cpu.set_variable_mode(maxdecoy.VARMOD_REFERENCE)
    assert n_found > 20 and n_exp > 20


def test_terminal_weight_examples(cpu):
    """Hand-checked weights: the terminal delta counts once, at its end, whatever the letter count."""
    pm = pyref.Mods((K_CTERM_FIX, M_NTERM_FIX), 0)
    base = pyref.sequence_weight
    assert pyref.SlotPeptide(pm, "MKAKMK").w == base("MKAKMK") + K_CTERM_FIX.mono_mass_int + M_NTERM_FIX.mono_mass_int
    assert pyref.SlotPeptide(pm, "KMAKMA").w == base("KMAKMA")
    assert pyref.SlotPeptide(pm, "K").w == base("K") + K_CTERM_FIX.mono_mass_int
    assert pyref.SlotPeptide(pm, "M").w == base("M") + M_NTERM_FIX.mono_mass_int
    pv = pyref.Mods((S_NTERM_FIX, S_NTERM_VAR), 2)
    sp = pyref.SlotPeptide(pv, "SAS")
    assert not sp.set_variable(0) and not sp.set_variable(2) and sp.w == base("SAS") + S_NTERM_FIX.mono_mass_int


def test_terminal_header_summary():
    """ModRes summary (modified_peptide.rs:606-659): side-chain, N-terminus and C-terminus slots, sorted by accession|name."""
    from maxdecoy import outputs
    mods = [K_CTERM_FIX, M_NTERM_FIX, synth.OXM]
    assert outputs.modification_summary("MKAKMK", mods, 1 << 4) == "(1|unimod:1|Acetyl)(1|unimod:35|Oxidation)(1|x:259|Label:13C(6)15N(2))"
    assert outputs.modification_summary("MKAKMK", mods, 1) == "(1|unimod:1|Acetyl)(1|unimod:35|Oxidation)(1|x:259|Label:13C(6)15N(2))"
    assert outputs.modification_summary("AKAKMA", mods, 0) == ""
    assert outputs.modification_summary("SAS", [S_NTERM_FIX, S_NTERM_VAR], 0) == "(1|x:3|Acetyl)"


@pytest.mark.parametrize("mods,nvar", MOD_SETS, ids=IDS)
def test_terminal_scores_and_decoys(cpu, mods, nvar):
    """Scores of targets and decoys against the dense Python table with the terminal masses on the end residues; every
    random / permuted decoy's slot-model weight lies in its window and equals the reported weight."""
    _setup(cpu, 120, mods, nvar)
    peptides = _table(cpu)
    seqs_t = [p[0] for p in peptides]
    pm = pyref.Mods(mods, nvar)
    sp = _spectra_for(cpu, peptides, mods, nvar, 8, seed=5)
    w = 20000
    total_dec = 0
    for mode in (maxdecoy.DECOY_REFERENCE_RANDOM, maxdecoy.DECOY_PERMUTE_TARGET):
        prm = SearchParams(10, 10, fragment_tolerance=0.02, n_decoys=12, decoy_mode=mode, seed=2, top_k=3, keep_decoys=True, abs_lower_uda=3_000_000,
                           abs_upper_uda=3_000_000)
        psms, st, scores, off = cpu.identify(sp, prm, want_all_scores=True)
        dec = cpu.last_decoys()
        dseq = wl.decoy_strings(dec)
        pre = [(P, P - 3_000_000, P + 3_000_000, z, sid) for (P, lo, hi, z, sid) in wl.precursors_of(cpu, sp)]
        cand = cpu.candidates(pre)
        n_dec = 0
        for s, (P, lo, hi, z, sid) in enumerate(pre):
            p0, p1 = int(sp.peak_off[s]), int(sp.peak_off[s + 1])
            T = pyref.xcorr_table(sp.peak_mz[p0:p1], sp.peak_intensity[p0:p1], P, w)
            want = []
            for i in range(int(cand["off"][s]), int(cand["off"][s + 1])):
                want.append(pyref.score(pm, T, seqs_t[int(cand["peptide_id"][i]) - 1], int(cand["var_mask"][i]), z, w))
            for i in range(int(dec["off"][s]), int(dec["off"][s + 1])):
                want.append(pyref.score(pm, T, dseq[i], int(dec["var_mask"][i]), z, w))
                mp = pyref.SlotPeptide(pm, dseq[i])
                m = int(dec["var_mask"][i])
                for j in range(len(dseq[i])):
                    if (m >> j) & 1:
                        assert mp.set_variable(j), (dseq[i], j)
                assert mp.w == int(dec["mod_weight"][i]) and lo <= mp.w <= hi
                assert dseq[i] not in seqs_t
                n_dec += 1
            assert scores[int(off[s]):int(off[s + 1])].tolist() == want, s
        assert n_dec > 0 or mode == maxdecoy.DECOY_PERMUTE_TARGET, mode
        total_dec += n_dec
    assert total_dec > 0


def _expanded_bruteforce(pm, seqs, lo, hi):
    """Every peptide x every subset (<= nvar) of the positions that can take their letter's variable modification, slot by slot."""
    import itertools
    want = set()
    span = pm.nvar * max([abs(v) for v in pm.var.values()] + [0])
    for p, q in enumerate(seqs):
        base = pyref.SlotPeptide(pm, q)
        if base.w - span > hi or base.w + span < lo:
            continue
        elig = []
        for j, c in enumerate(q):
            if c in pm.var:
                t = pyref.SlotPeptide(pm, q)
                if t.set_variable(j):
                    elig.append(j)
        for n in range(0, pm.nvar + 1):
            for sub in itertools.combinations(elig, n):
                w = base.w + sum(pm.var[q[j]] for j in sub)
                if lo <= w <= hi:
                    want.add((p, sum(1 << j for j in sub), w))
    return want


@pytest.mark.parametrize("mods,nvar", MOD_SETS[1:], ids=IDS[1:])
def test_terminal_expanded_mode_equals_bruteforce(cpu, mods, nvar):
    """MD_VARMOD_EXPANDED with terminal modifications: every placement of <= nvar variable modifications on the positions
    where their slot exists and is free, against the brute force over all peptides and subsets."""
    _setup(cpu, 100, mods, nvar)
    peptides = _table(cpu)
    seqs = [p[0] for p in peptides]
    pm = pyref.Mods(mods, nvar)
    pre = _precursors(pm, peptides, seed=21, n=16, narrow=10)
    pre = [(P, lo, hi, z, sid) if hi - lo < 1_000_000 else (P, P - 1_500_000, P + 1_500_000, z, sid) for (P, lo, hi, z, sid) in pre]
    cpu.set_variable_mode(maxdecoy.VARMOD_EXPANDED)
    try:
        cpu.index_build()
        exp = cpu.candidates(pre)
        n_var = 0
        for s, (P, lo, hi, z, sid) in enumerate(pre):
            a, b = int(exp["off"][s]), int(exp["off"][s + 1])
            got = [(int(exp["peptide_id"][i]) - 1, int(exp["var_mask"][i]), int(exp["mod_weight"][i])) for i in range(a, b)]
            assert len(set(got)) == len(got)
            assert set(got) == _expanded_bruteforce(pm, seqs, lo, hi), s
            n_var += sum(1 for g in got if g[1])
        assert int(exp["off"][-1]) > 0
        if mods != MOD_SETS[4][0]:
            assert n_var > 0
    finally:
        cpu.set_variable_mode(maxdecoy.VARMOD_REFERENCE)
        cpu.index_build()


def test_terminal_unsupported_modes(cpu):
    _setup(cpu, 40, (K_CTERM_FIX,), 0)
    pre = [(900_400_000, 900_300_000, 900_500_000, 2, 0)]
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.generate_decoys(pre, 5, maxdecoy.DECOY_EXHAUSTIVE, seed=0)


# ------------------------------------------------------------------------------------------------ GPU: CUDA vs oracle
@pytest.fixture(scope="module")
def gpu():
    e = maxdecoy.Engine()
    assert e.backend == "cuda-sm100a"
    yield e
    e.close()


def _equal(a, b):
    for k in a.keys():
        assert a[k].shape == b[k].shape, k
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("mods,nvar", MOD_SETS, ids=IDS)
def test_terminal_gpu_bit_exact(gpu, cpu, mods, nvar):
    for e in (gpu, cpu):
        _setup(e, 300, mods, nvar)
    n = cpu.index_stats()["n_peptides"]
    pg, kg = gpu.index_export(0, n)
    pc, kc = cpu.index_export(0, n)
    assert np.array_equal(kg, kc) and np.array_equal(pg, pc)
    peptides = _table(cpu)
    pm = pyref.Mods(mods, nvar)
    pre = _precursors(pm, peptides, seed=3, n=48, narrow=24)
    cg, cc = gpu.candidates(pre), cpu.candidates(pre)
    _equal(cg, cc)
    assert len(cc["peptide_id"]) > 0
    for mode, nd in ((maxdecoy.DECOY_REFERENCE_RANDOM, 60), (maxdecoy.DECOY_PERMUTE_TARGET, 20)):
        dg = gpu.generate_decoys(pre[:32], nd, mode, seed=9)
        dc = cpu.generate_decoys(pre[:32], nd, mode, seed=9)
        _equal(dg, dc)
        assert len(dc["attempt"]) > 0
    sp = _spectra_for(cpu, peptides, mods, nvar, 40, seed=8)
    for mode, nd, topk in ((maxdecoy.DECOY_REFERENCE_RANDOM, 40, 5), (maxdecoy.DECOY_PERMUTE_TARGET, 10, 12)):
        prm = SearchParams(10, 10, n_decoys=nd, decoy_mode=mode, seed=3, top_k=topk, abs_lower_uda=2_000_000, abs_upper_uda=2_000_000)
        rg = gpu.identify(sp, prm, want_all_scores=True)
        rc = cpu.identify(sp, prm, want_all_scores=True)
        assert np.array_equal(rg[3], rc[3]) and np.array_equal(rg[2], rc[2])
        for f in ("spectrum_id", "rank", "is_decoy", "charge", "candidate", "var_mask", "mod_weight", "raw_score", "n_targets", "n_decoys"):
            assert np.array_equal(rg[0][f], rc[0][f]), f
        assert np.allclose(rg[0]["score"], rc[0]["score"], rtol=1e-5, atol=0)
        assert int(rc[0]["n_targets"].sum()) > 0
    # the generating peptides are found: narrow windows this time
    prm = SearchParams(10, 10, n_decoys=0, top_k=1)
    rg, rc = gpu.identify(sp, prm), cpu.identify(sp, prm)
    assert np.array_equal(rg[0]["raw_score"], rc[0]["raw_score"]) and np.array_equal(rg[0]["candidate"], rc[0]["candidate"])


@pytest.mark.gpu
@pytest.mark.parametrize("mods,nvar", MOD_SETS[1:], ids=IDS[1:])
def test_terminal_expanded_gpu_bit_exact(gpu, cpu, mods, nvar):
    for e in (gpu, cpu):
        e.digest(list(wl.proteins(300)), 2, 5, 50)
        e.set_modifications(list(mods), nvar)
        e.set_variable_mode(maxdecoy.VARMOD_EXPANDED)
        e.index_build()
    try:
        peptides = _table(cpu)
        pm = pyref.Mods(mods, nvar)
        pre = _precursors(pm, peptides, seed=4, n=48, narrow=24)
        _equal(gpu.candidates(pre), cpu.candidates(pre))
        sp = _spectra_for(cpu, peptides, mods, nvar, 32, seed=9)
        prm = SearchParams(10, 10, n_decoys=20, seed=5, top_k=5, abs_lower_uda=1_500_000, abs_upper_uda=1_500_000)
        rg = gpu.identify(sp, prm, want_all_scores=True)
        rc = cpu.identify(sp, prm, want_all_scores=True)
        assert np.array_equal(rg[3], rc[3]) and np.array_equal(rg[2], rc[2])
        for f in ("rank", "is_decoy", "candidate", "var_mask", "mod_weight", "raw_score", "n_targets", "n_decoys"):
            assert np.array_equal(rg[0][f], rc[0][f]), f
        assert int(rc[0]["n_targets"].sum()) > 0
    finally:
        for e in (gpu, cpu):
            e.set_variable_mode(maxdecoy.VARMOD_REFERENCE)


@pytest.mark.gpu
def test_terminal_gpu_unsupported_modes(gpu):
    _setup(gpu, 40, (K_CTERM_FIX,), 0)
    pre = [(900_400_000, 900_300_000, 900_500_000, 2, 0)]
    with pytest.raises(maxdecoy.MaxDecoyError):
        gpu.generate_decoys(pre, 5, maxdecoy.DECOY_EXHAUSTIVE, seed=0)
