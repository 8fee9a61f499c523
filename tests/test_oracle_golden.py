"""Pins the CPU oracle (oracle/maxdecoy_oracle.cpp) before anything is compared against it:
  * the reference's own golden vector and known-answer values,
  * a second, deliberately naive pure-Python restatement of the reference text (tests/pyref.py) on small cases,
  * the properties the reference guarantees for its (unseeded, irreproducible) random decoys.
CPU only: runs in the `-m "not gpu"` suite."""
import numpy as np
import pytest

import maxdecoy
from maxdecoy import SearchParams, synth
from oracle_lib import oracle_engine
import pyref
import workloads as wl


@pytest.fixture(scope="module")
def cpu():
    e = oracle_engine(4)
    assert e.backend == "cpu-oracle"
    yield e
    e.close()


# ---------------------------------------------------------------------------------------------- masses
# int(mono * 1e6) truncating; E and K differ from round-to-nearest by -1
TABLE = {"A": 71037110, "B": 114534950, "R": 156101110, "N": 114042930, "D": 115026940, "C": 103009190,
         "E": 129042589, "Q": 128058580, "G": 57021460, "H": 137058910, "I": 113084060, "L": 113084060,
         "J": 113084060, "K": 128094959, "M": 131040490, "F": 147068410, "P": 97052760, "O": 109052800,
         "S": 87032030, "T": 101047680, "U": 150953630, "V": 99068410, "W": 186079310, "X": 0,
         "Y": 163063330, "Z": 128550590}


def test_mass_table_truncation_quirk(cpu):
    for c, v in TABLE.items():
        assert cpu.residue_mass(c) == v, c
        assert pyref.residue_mass(c) == v, c
    for c in "?*a1 ":
        assert cpu.residue_mass(c) == 0        # unknown -> X -> 0 (amino_acid.rs:115)


@pytest.mark.parametrize("seq,weight", [("VVGTVK", 601379894), ("DHWVHVJVPMGFVJGCYJDR", 2355165605),
                                        ("MLLRAG", 659378855), ("MSLREK", 762405803), ("", 18010565)])
def test_sequence_weight_known_answers(cpu, seq, weight):
    assert cpu.get_sequence_weight(seq) == weight
    assert pyref.sequence_weight(seq) == weight


def test_modification_mass_conversion():
    assert synth.CAM.mono_mass_int == 57021464 and synth.OXM.mono_mass_int == 15994915


# ---------------------------------------------------------------------------------------------- digest
def test_digest_golden_p77377(cpu):
    """The reference's only known-answer test: models/enzyms/tests/digest_enzym.rs:13-92."""
    g = wl.p77377()
    assert len(g["sequence"]) == 492 and len(g["peptides"]) == 71
    n = cpu.digest([g["sequence"]], **g["params"])
    assert n == 71
    t = cpu.peptides()
    got = cpu.sequences_of(t)
    want = {pyref.generalize(s) for s in g["peptides"]}
    assert set(got) == want and len(got) == 71
    for s, w, c in zip(got, t["weight"], t["counts"]):
        assert int(w) == pyref.sequence_weight(s)
        assert list(c) == pyref.counts21(s)
    # peptides/tests/peptide.rs:17-21
    vv = pyref.counts21("VVGTVK")
    assert vv[pyref.ALPHABET.index("V")] == 3


@pytest.mark.parametrize("mc,min_len,max_len", [(0, 1, 60), (1, 5, 50), (2, 5, 50), (4, 6, 30)])
def test_digest_matches_python_restatement(cpu, mc, min_len, max_len):
    prots = list(wl.proteins(60)) + ["", "K", "KP", "RRRR", "AKPAKAR", "MKKPKKR", "ABZXUOK", "ILILIK"]
    cpu.digest(prots, mc, min_len, max_len)
    t = cpu.peptides()
    want = {}
    for pi, p in enumerate(prots):
        for s, m in pyref.digest(p, mc, min_len, max_len):
            e = want.setdefault(s, [m, set()])
            e[0] = min(e[0], m)
            e[1].add(pi)
    got = cpu.sequences_of(t)
    assert len(got) == len(set(got))                      # UNIQUE(aa_sequence, weight), schema.sql:41
    assert set(got) == set(want)
    for k, s in enumerate(got):
        assert int(t["missed_cleavages"][k]) == want[s][0]
        assert int(t["weight"][k]) == pyref.sequence_weight(s)
        assert list(t["counts"][k]) == pyref.counts21(s)
        a0, a1 = int(t["assoc_off"][k]), int(t["assoc_off"][k + 1])
        assert set(t["assoc_protein"][a0:a1].tolist()) == want[s][1]
    # canonical order: weight ascending
    assert np.all(np.diff(t["weight"]) >= 0)


def test_digest_argument_errors(cpu):
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.digest(["MKR"], 2, 5, 61)                     # max_len > 60 (tasks/digestion.rs:101)
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.digest(["MKR"], 61, 5, 50)                    # missed cleavages > 60 (tasks/digestion.rs:74)
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.digest(["MKR"], 2, 0, 50)


# ---------------------------------------------------------------------------------------------- window math
def test_precursor_window_matches_python_f64(cpu):
    rng = np.random.default_rng(5)
    for _ in range(3000):
        mz = float(rng.uniform(150.0, 2500.0))
        z = int(rng.integers(1, 7))
        lppm, uppm = int(rng.integers(0, 50)), int(rng.integers(0, 50))
        assert cpu.precursor_window(mz, z, lppm, uppm) == pyref.precursor_window(mz, z, lppm, uppm)
    P, lo, hi = cpu.precursor_window(500.0, 2, 10, 10)
    assert lo <= P <= hi and hi - lo == pytest.approx(2 * 2 * 500.0 * 10, abs=3)


# ---------------------------------------------------------------------------------------------- lookup
def _table(cpu):
    t = cpu.peptides()
    seqs = cpu.sequences_of(t)
    return [(s, int(w), list(c)) for s, w, c in zip(seqs, t["weight"], t["counts"])]


@pytest.mark.parametrize("mods,nvar", [((synth.CAM,), 0), ((synth.CAM, synth.OXM), 3), ((synth.OXM,), 2)])
def test_candidates_equal_literal_sql_fanout(cpu, mods, nvar):
    """The oracle's single W* window + filter against the reference's procedure run literally: every fan-out
    query (identification.rs:374-403) as a table scan, then ModifiedPeptide (identification.rs:242-257)."""
    cpu.digest(list(wl.proteins(40)), 2, 5, 50)
    cpu.set_modifications(list(mods), nvar)
    cpu.index_build()
    peptides = _table(cpu)
    pm = pyref.Mods(mods, nvar)
    sp, _ = synth.synthetic_spectra(list(wl.proteins(40)), 14, 2, mods=tuple(mods), seed=3)
    pre = wl.precursors_of(cpu, sp)
    # plus wide windows, where several variable-modification counts can hit
    pre += [(p[0], p[0] - 20_000_000, p[0] + 20_000_000, p[3], 100 + i) for i, p in enumerate(pre[:4])]
    got = cpu.candidates(pre)
    n_found = 0
    for s, (P, lo, hi, z, sid) in enumerate(pre):
        want = pyref.candidates_sql(pm, peptides, P, lo, hi)
        a, b = int(got["off"][s]), int(got["off"][s + 1])
        have = {int(got["peptide_id"][i]) - 1: (int(got["mod_weight"][i]), int(got["var_mask"][i])) for i in range(a, b)}
        assert have == want, s
        n_found += len(want)
    assert n_found > 0


def test_expanded_variable_mode_equals_bruteforce(cpu):
    """MD_VARMOD_EXPANDED: every subset of up to nvar variable-modifiable residues whose weight lies in
    the window is a candidate -- against a brute force over all peptides and all subsets; and the reference mode's
    candidates are a subset of it wherever the reference finds anything (same peptide, same placement, same weight)."""
    import itertools
    from maxdecoy import Modification
    var_s = Modification("x:21", "Phospho", "A", False, "S", 79.966331)
    mods, nvar = (synth.CAM, synth.OXM, var_s), 2
    prots = list(wl.proteins(120))
    cpu.digest(prots, 2, 5, 50)
    cpu.set_modifications(list(mods), nvar)
    cpu.set_variable_mode(maxdecoy.VARMOD_REFERENCE)
    cpu.index_build()
    seqs = cpu.sequences_of(cpu.peptides())
    pm = pyref.Mods(mods, nvar)
    rng = np.random.default_rng(5)
    pre = []
    for i in range(14):                                     # precursors of partially modified peptides
        q = seqs[int(rng.integers(len(seqs)))]
        elig = [j for j, c in enumerate(q) if c in pm.var and c not in pm.fix]
        pick = [j for j in elig if rng.random() < 0.5][:nvar]
        m = pyref.sequence_weight(q) + sum(pm.fix.get(c, 0) for c in q) + sum(pm.var[q[j]] for j in pick)
        tol = m // 100_000
        pre.append((m, m - tol, m + tol, 2, i))
    ref = cpu.candidates(pre)
    cpu.set_variable_mode(maxdecoy.VARMOD_EXPANDED)
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.candidates(pre)                                 # the mode change invalidated the index
    cpu.index_build()
    exp = cpu.candidates(pre)
    wfix = [pyref.sequence_weight(q) + sum(pm.fix.get(c, 0) for c in q) for q in seqs]
    found_partial = 0
    for s, (P, lo, hi, z, sid) in enumerate(pre):
        a, b = int(exp["off"][s]), int(exp["off"][s + 1])
        got = [(int(exp["peptide_id"][i]) - 1, int(exp["var_mask"][i]), int(exp["mod_weight"][i])) for i in range(a, b)]
        assert len(set(got)) == len(got)
        want = set()
        for p, q in enumerate(seqs):
            if wfix[p] > hi or wfix[p] + nvar * 80_000_000 < lo:
                continue
            elig = [j for j, c in enumerate(q) if c in pm.var and c not in pm.fix]
            for n in range(0, nvar + 1):
                for sub in itertools.combinations(elig, n):
                    w = wfix[p] + sum(pm.var[q[j]] for j in sub)
                    if lo <= w <= hi:
                        want.add((p, sum(1 << j for j in sub), w))
        assert set(got) == want
        assert len(want) >= 1                               # the generating form itself
        found_partial += any(0 < bin(m).count("1") < sum(1 for c in seqs[p] if c in pm.var and c not in pm.fix) for p, m, w in got)
        ra, rb = int(ref["off"][s]), int(ref["off"][s + 1])
        for i in range(ra, rb):
            assert (int(ref["peptide_id"][i]) - 1, int(ref["var_mask"][i]), int(ref["mod_weight"][i])) in want
    assert found_partial > 0                                # forms the reference mode cannot find
    cpu.set_variable_mode(maxdecoy.VARMOD_REFERENCE)
    cpu.index_build()


def test_no_modifiable_letter_means_no_targets(cpu):
    """identification.rs:375-379: with an empty modification list the recursion emits no query."""
    cpu.digest(list(wl.proteins(40)), 2, 5, 50)
    cpu.set_modifications([], 0)
    cpu.index_build()
    sp, _ = wl.spectra(40, 10, 2)
    got = cpu.candidates(wl.precursors_of(cpu, sp))
    assert int(got["off"][-1]) == 0


def test_nchoosek_order():
    # n_choose_k.rs: 4 choose 2 -> 1100, 1010, 1001, 0110, 0101, 0011 (MSB = first item)
    assert list(pyref.n_choose_k_masks(4, 2)) == [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]


def test_window_search_bounds(cpu):
    cpu.digest(list(wl.proteins(40)), 2, 5, 50)
    cpu.set_modifications([synth.CAM], 0)
    cpu.index_build()
    st = cpu.index_stats()
    _, key = cpu.index_export(0, st["n_peptides"])
    assert np.all(np.diff(key) >= 0) and int(key[0]) == st["min_key"] and int(key[-1]) == st["max_key"]
    lo = np.array([0, int(key[10]), int(key[10]) + 1, int(key[-1]) + 1, 5], dtype=np.int64)
    hi = np.array([10**12, int(key[10]), int(key[10]), 10**12, 4], dtype=np.int64)
    b, e = cpu.window_search(lo, hi)
    assert np.array_equal(b, np.searchsorted(key, lo, "left"))
    assert np.array_equal(e, np.maximum(np.searchsorted(key, hi, "right"), b))


# ---------------------------------------------------------------------------------------------- decoys
@pytest.mark.parametrize("mods,nvar", [((synth.CAM,), 0), ((synth.CAM, synth.OXM), 3)])
def test_random_decoy_properties(cpu, mods, nvar):
    """What the reference guarantees for its random decoys: alphabet, modified mass inside the
    window, not a target peptide, unique per spectrum, length <= 60, var-mod count <= NVAR on var letters only."""
    cpu.digest(list(wl.proteins(150)), 2, 5, 50)
    cpu.set_modifications(list(mods), nvar)
    cpu.index_build()
    targets = set(cpu.sequences_of(cpu.peptides()))
    sp, _ = wl.spectra(150, 12, 2, with_ox=len(mods) > 1)
    pre = wl.precursors_of(cpu, sp)
    d = cpu.generate_decoys(pre, 120, maxdecoy.DECOY_REFERENCE_RANDOM, seed=42)
    seqs = wl.decoy_strings(d)
    pm = pyref.Mods(mods, nvar)
    total = 0
    for s, (P, lo, hi, z, sid) in enumerate(pre):
        a, b = int(d["off"][s]), int(d["off"][s + 1])
        assert b - a <= 120
        mine = seqs[a:b]
        assert len(set(mine)) == len(mine)
        for i in range(a, b):
            q, mask = seqs[i], int(d["var_mask"][i])
            assert 1 <= len(q) <= 60 and set(q) <= set(pyref.ALPHABET)
            assert q not in targets
            assert int(d["weight"][i]) == pyref.sequence_weight(q)           # Decoy::new: unmodified weight
            w = pyref.sequence_weight(q) + sum(pm.fix.get(c, 0) for c in q)
            npos = 0
            for k, c in enumerate(q):
                if (mask >> k) & 1:
                    assert c in pm.var and c not in pm.fix
                    w += pm.var[c]
                    npos += 1
            assert npos <= nvar
            assert w == int(d["mod_weight"][i]) and lo <= w <= hi
        assert np.all(np.diff(d["attempt"][a:b].astype(np.int64)) > 0)       # attempt order
        total += b - a
    assert total > 0
    # seeded: same call -> same decoys; a subset of the spectra -> the same decoys for those spectra
    d2 = cpu.generate_decoys(pre, 120, maxdecoy.DECOY_REFERENCE_RANDOM, seed=42)
    assert all(np.array_equal(d[k], d2[k]) for k in d)
    d3 = cpu.generate_decoys(pre[5:7], 120, maxdecoy.DECOY_REFERENCE_RANDOM, seed=42)
    assert wl.decoy_strings(d3) == seqs[int(d["off"][5]):int(d["off"][7])]
    d4 = cpu.generate_decoys(pre, 120, maxdecoy.DECOY_REFERENCE_RANDOM, seed=43)
    assert wl.decoy_strings(d4) != seqs


@pytest.mark.parametrize("mods,nvar", [((synth.CAM,), 0), ((synth.CAM, synth.OXM), 3)])
def test_stored_decoys_are_reused_before_new_ones(cpu, mods, nvar):
    """tasks/identification.rs:259-283: decoys of the `decoys` table that the target queries retrieve and the ModifiedPeptide
    filter accepts are taken first; only the remainder is generated.  Checked against the literal query-by-query Python
    execution (pyref.candidates_sql) over the stored sequences."""
    cpu.digest(list(wl.proteins(150)), 2, 5, 50)
    cpu.set_modifications(list(mods), nvar)
    cpu.index_build()
    cpu.set_decoy_store([])
    sp, _ = wl.spectra(150, 10, 2, with_ox=len(mods) > 1)
    pre = wl.precursors_of(cpu, sp)
    fresh = cpu.generate_decoys(pre, 60, maxdecoy.DECOY_REFERENCE_RANDOM, seed=9)
    seqs = wl.decoy_strings(fresh)
    assert not np.any(fresh["attempt"] == maxdecoy.DECOY_STORED)
    # the store: every second decoy of the first run, twice (duplicates are dropped), plus sequences that match nothing
    store = seqs[::2] + seqs[::2] + ["GGGGG", "WWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWW"]
    cpu.set_decoy_store(store)
    pm = pyref.Mods(mods, nvar)
    uniq = sorted(set(store))
    table = [(q, pyref.sequence_weight(q), pyref.counts21(q)) for q in uniq]
    got = cpu.generate_decoys(pre, 60, maxdecoy.DECOY_REFERENCE_RANDOM, seed=9)
    gs = wl.decoy_strings(got)
    for s, (P, lo, hi, z, sid) in enumerate(pre):
        a, b = int(got["off"][s]), int(got["off"][s + 1])
        mine, att = gs[a:b], got["attempt"][a:b]
        assert len(set(mine)) == len(mine)
        stored = [i for i in range(b - a) if att[i] == maxdecoy.DECOY_STORED]
        assert stored == list(range(len(stored)))                          # stored decoys come first
        want = {uniq[i]: wm for i, wm in pyref.candidates_sql(pm, table, P, lo, hi).items()}
        assert len(stored) == min(60, len(want))
        for i in stored:
            assert mine[i] in want
            assert (int(got["mod_weight"][a + i]), int(got["var_mask"][a + i])) == want[mine[i]]
            assert int(got["weight"][a + i]) == pyref.sequence_weight(mine[i])
        if len(want) <= 60:
            assert {mine[i] for i in stored} == set(want)
        # the generated remainder: the attempts of the first run in order, minus the sequences already taken from the store
        fa, fb = int(fresh["off"][s]), int(fresh["off"][s + 1])
        rest = [q for q in seqs[fa:fb] if q not in set(mine[:len(stored)])]
        assert mine[len(stored):] == rest[:len(mine) - len(stored)]
    assert np.any(got["attempt"] == maxdecoy.DECOY_STORED)
    # changing the modifications re-indexes the store; clearing it restores the first run
    cpu.set_decoy_store([])
    again = cpu.generate_decoys(pre, 60, maxdecoy.DECOY_REFERENCE_RANDOM, seed=9)
    assert all(np.array_equal(fresh[k], again[k]) for k in fresh)
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.set_decoy_store(["PEPTIDEB"])                                  # B is not in the decoy alphabet
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.set_decoy_store([""])


def test_permuted_target_decoys(cpu):
    """vary_targets (decoy_generator.rs:265-296): a shuffled target, same composition, not a peptide."""
    cpu.digest(list(wl.proteins(150)), 2, 5, 50)
    cpu.set_modifications([synth.CAM], 0)
    cpu.index_build()
    targets = set(cpu.sequences_of(cpu.peptides()))
    sp, _ = wl.spectra(150, 12, 2)
    pre = wl.precursors_of(cpu, sp)
    cand = cpu.candidates(pre)
    seqs_t = cpu.sequences_of(cpu.peptides())
    d = cpu.generate_decoys(pre, 20, maxdecoy.DECOY_PERMUTE_TARGET, seed=9)
    seqs = wl.decoy_strings(d)
    for s in range(len(pre)):
        comp = {"".join(sorted(seqs_t[int(p) - 1])) for p in cand["peptide_id"][int(cand["off"][s]):int(cand["off"][s + 1])]}
        for i in range(int(d["off"][s]), int(d["off"][s + 1])):
            assert "".join(sorted(seqs[i])) in comp and seqs[i] not in targets
            assert pre[s][1] <= int(d["mod_weight"][i]) <= pre[s][2]
    assert len(seqs) > 0


@pytest.mark.parametrize("lo,hi,max_len", [(18010565 + 57021460, 18010565 + 3 * 57021460, 3), (260_000_000, 260_400_000, 4),
                                           (288_000_000, 290_300_000, 4)])
def test_exhaustive_decoys_equal_bruteforce(cpu, lo, hi, max_len):
    """Exhaustive-enumeration mode against a brute force over every sequence (21^L) for tiny precursors."""
    cpu.digest(["GGK", "AGGR", "GAK"] + list(wl.proteins(5)), 2, 2, 50)
    cpu.set_modifications([synth.CAM], 0)
    cpu.index_build()
    targets = set(cpu.sequences_of(cpu.peptides()))
    pm = pyref.Mods((synth.CAM,), 0)
    allseq = pyref.exhaustive_bruteforce(pm, lo, hi, max_len)
    # nothing longer than max_len can fit: the lightest residue is G
    assert pyref.to_int(pyref.H2O) + (max_len + 1) * TABLE["G"] > hi
    want = [(q, w, k) for k, (q, w) in enumerate(allseq) if q not in targets]
    n = 10**6
    d = cpu.generate_decoys([((lo + hi) // 2, lo, hi, 2, 0)], n, maxdecoy.DECOY_EXHAUSTIVE, seed=0)
    got = list(zip(wl.decoy_strings(d), d["mod_weight"].tolist(), d["attempt"].tolist()))
    assert got == want and len(want) > 0
    # truncation keeps the canonical prefix
    d5 = cpu.generate_decoys([((lo + hi) // 2, lo, hi, 2, 0)], 5, maxdecoy.DECOY_EXHAUSTIVE, seed=0)
    assert wl.decoy_strings(d5) == [q for q, _, _ in want[:5]]


# ---------------------------------------------------------------------------------------------- scoring
@pytest.mark.parametrize("mods,nvar,tol", [((synth.CAM,), 0, 0.02), ((synth.CAM, synth.OXM), 3, 0.02), ((synth.CAM,), 0, 1.0005)])
def test_scores_equal_dense_python_xcorr(cpu, mods, nvar, tol):
    """The oracle's sparse prefix-sum xcorr against a dense table built peak by peak (pyref.xcorr_table)."""
    cpu.digest(list(wl.proteins(150)), 2, 5, 50)
    cpu.set_modifications(list(mods), nvar)
    cpu.index_build()
    seqs_t = cpu.sequences_of(cpu.peptides())
    sp, truth = wl.spectra(150, 10, 2, with_ox=len(mods) > 1)
    prm = SearchParams(10, 10, fragment_tolerance=tol, n_decoys=8, seed=1, top_k=3, keep_decoys=True)
    psms, st, scores, off = cpu.identify(sp, prm, want_all_scores=True)
    dec = cpu.last_decoys()
    dseq = wl.decoy_strings(dec)
    pre = wl.precursors_of(cpu, sp)
    cand = cpu.candidates(pre)
    pm = pyref.Mods(mods, nvar)
    w = int(round(tol * 1e6))
    hits = 0
    for s, (P, lo, hi, z, sid) in enumerate(pre):
        p0, p1 = int(sp.peak_off[s]), int(sp.peak_off[s + 1])
        T = pyref.xcorr_table(sp.peak_mz[p0:p1], sp.peak_intensity[p0:p1], P, w)
        want = []
        for i in range(int(cand["off"][s]), int(cand["off"][s + 1])):
            want.append(pyref.score(pm, T, seqs_t[int(cand["peptide_id"][i]) - 1], int(cand["var_mask"][i]), z, w))
        for i in range(int(dec["off"][s]), int(dec["off"][s + 1])):
            want.append(pyref.score(pm, T, dseq[i], int(dec["var_mask"][i]), z, w))
        got = scores[int(off[s]):int(off[s + 1])].tolist()
        assert got == want, s
        # top-k rows: raw score descending, candidate ordinal ascending
        order = sorted(range(len(want)), key=lambda i: (-want[i], i))[:3]
        for r, i in enumerate(order):
            assert int(psms["rank"][s, r]) == r + 1 and int(psms["raw_score"][s, r]) == want[i]
            assert psms["score"][s, r] == np.float32(0.005 * want[i] / (150.0 * 65536.0))
        # the generating peptide of a database spectrum should win
        nt = int(cand["off"][s + 1] - cand["off"][s])
        if nt and order and order[0] < nt and seqs_t[int(cand["peptide_id"][int(cand["off"][s]) + order[0]]) - 1] == truth[s][0]:
            hits += 1
    assert hits >= 5


def test_unscorable_spectra_and_empty_batch(cpu):
    cpu.digest(list(wl.proteins(40)), 2, 5, 50)
    cpu.set_modifications([synth.CAM], 0)
    cpu.index_build()
    sp, _ = wl.spectra(40, 4, 2)
    few = maxdecoy.Spectra(sp.precursor_mz[:2], sp.charge[:2], np.array([0, 3, 3], dtype=np.uint64), sp.peak_mz[:3], sp.peak_intensity[:3])
    psms, st = cpu.identify(few, SearchParams(10, 10, n_decoys=5, top_k=2))
    assert np.all(psms["rank"] == 0) and st["n_spectra"] == 2          # fewer than minimum_peaks -> not scored
    empty = maxdecoy.Spectra(np.zeros(0), np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.uint64), np.zeros(0), np.zeros(0, dtype=np.float32))
    psms, st = cpu.identify(empty, SearchParams(10, 10, n_decoys=5, top_k=2))
    assert psms.shape == (0, 2) and st["n_spectra"] == 0
    bad = maxdecoy.Spectra(sp.precursor_mz[:1], sp.charge[:1], np.array([0, 3], dtype=np.uint64), np.array([300.0, 200.0, 400.0]), np.ones(3, dtype=np.float32))
    with pytest.raises(maxdecoy.MaxDecoyError):
        cpu.identify(bad, SearchParams(10, 10))                       # peaks must be sorted by m/z


def test_substitution_map(cpu):
    """get_one_amino_acid_substitute_map (decoy_generator.rs:301-324)."""
    cpu.set_modifications([synth.CAM], 0)
    m = cpu.substitution_map()
    a = pyref.ALPHABET
    mp = [TABLE[c] + (57021464 if c == "C" else 0) for c in a]
    for i in range(21):
        for j in range(21):
            assert int(m[i, j]) == mp[j] - mp[i]
