"""Host-side logic and the C-ABI boundary, without a GPU: the CUDA library loads and exports every symbol
include/maxdecoy.h declares (no compute calls), refuses to run without a device (no CPU fallback), and the
Python mirror of the reference's interface (modification CSV, FASTA/MGF readers, sharding) behaves."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import maxdecoy
from maxdecoy import _abi, parallel, synth
import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "maxdecoy.h")
CUDA_SO = os.path.join(ROOT, "max-decoy_b200", "csrc", "libmaxdecoy_cuda.so")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"MD_API\s+[\w\s\*]+?\b(md_\w+)\s*\(", text)))


def _ensure_cuda_lib():
    if not os.path.exists(CUDA_SO):
        subprocess.check_call(["make", "-C", os.path.dirname(CUDA_SO), "-j8", "-s"])
    return CUDA_SO


def test_header_and_ctypes_view_agree():
    syms = header_symbols()
    assert len(syms) >= 25
    assert set(syms) == set(_abi.SYMBOLS), set(syms) ^ set(_abi.SYMBOLS)


def test_rust_binding_declares_every_symbol():
    """ffi/maxdecoy_sys.rs (the uncompiled Rust view a maintainer of the reference would add) and the code block of
    INTEGRATION.md name exactly the header's entry points."""
    rs = open(os.path.join(ROOT, "ffi", "maxdecoy_sys.rs")).read()
    rs_syms = set(re.findall(r"pub fn (md_\w+)\s*\(", rs))
    assert rs_syms == set(header_symbols()), rs_syms ^ set(header_symbols())
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for s in header_symbols():
        assert s in doc, "INTEGRATION.md does not mention %s" % s


def test_cuda_library_exports_every_declared_symbol():
    lib = _abi.bind(_ensure_cuda_lib())            # AttributeError on a missing symbol
    out = subprocess.check_output(["nm", "-D", "--defined-only", CUDA_SO], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for s in header_symbols():
        assert s in exported, s
    leaked = {s for s in exported if not s.startswith("md_")}
    assert not leaked, "non-ABI symbols exported: %s" % sorted(leaked)[:5]
    assert lib.md_backend_name() == b"cuda-sm100a"


def test_cuda_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", _ensure_cuda_lib()], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_pure_entry_points_need_no_device():
    lib = _abi.bind(_ensure_cuda_lib())
    assert lib.md_residue_mass(ord("E")) == 129042589 and lib.md_residue_mass(ord("K")) == 128094959
    assert lib.md_sequence_weight(b"VVGTVK", 6) == 601379894
    P, lo, hi = C.c_int64(), C.c_int64(), C.c_int64()
    assert lib.md_precursor_window(500.0, 2, 10, 10, C.byref(P), C.byref(lo), C.byref(hi)) == 0
    import pyref
    assert (P.value, lo.value, hi.value) == pyref.precursor_window(500.0, 2, 10, 10)
    rng = np.random.default_rng(1)
    for _ in range(2000):
        mz, z = float(rng.uniform(150, 2500)), int(rng.integers(1, 7))
        lib.md_precursor_window(mz, z, 7, 13, C.byref(P), C.byref(lo), C.byref(hi))
        assert (P.value, lo.value, hi.value) == pyref.precursor_window(mz, z, 7, 13)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _abi.bind(_ensure_cuda_lib())
    cfg, h = _abi.md_config(0, 0), _abi.ctx_p()
    rc = lib.md_create(C.byref(cfg), C.byref(h))
    assert rc == -3 and not h.value                    # MD_ERR_DEVICE
    assert b"no CPU fallback" in lib.md_last_error(None)
    with pytest.raises(maxdecoy.MaxDecoyError):
        maxdecoy.Engine()


def test_product_never_touches_the_oracle():
    """Nothing under max-decoy_b200/ may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "max-decoy_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "libmaxdecoy_oracle" not in text and "oracle_lib" not in text, os.path.join(dp, f)
    out = subprocess.check_output(["ldd", _ensure_cuda_lib()], text=True)
    assert "oracle" not in out


# ------------------------------------------------------------------------------------------ host mirror
def test_modification_csv_loader(tmp_path):
    p = tmp_path / "mods.csv"
    p.write_text(synth.mods_csv_text([synth.CAM, synth.OXM]))
    mods = maxdecoy.Modification.create_from_csv_file(str(p))
    assert [(m.accession, m.name, m.position, m.is_fix, m.amino_acid, m.mono_mass_int) for m in mods] == [
        ("unimod:4", "Carbamidomethyl", "A", True, "C", 57021464), ("unimod:35", "Oxidation", "A", False, "M", 15994915)]
    p.write_text("h1,h2,h3,h4,h5,h6\nUNIMOD:4,Carbamidomethyl,A,1,C\n")
    with pytest.raises(ValueError):                     # modification.rs:59-61: "row has wrong length"
        maxdecoy.Modification.create_from_csv_file(str(p))


def test_fasta_round_trip_and_reference_quirk():
    prots = list(wl.proteins(30))
    hdr, seqs = synth.read_fasta(synth.fasta_text(prots))
    assert seqs == prots and all(h.startswith(">sp|P") for h in hdr)
    # lines before the first header are prepended to the first protein (fasta_digester.rs:75,113)
    _, seqs = synth.read_fasta("AAA\n>x\nKKK\n>y\nRRR\n")
    assert seqs == ["AAAKKK", "RRR"]


def test_mgf_round_trip():
    sp, _ = wl.spectra(40, 6, 2)
    back = synth.read_mgf(synth.mgf_text(sp))
    for k in ("precursor_mz", "charge", "peak_off", "peak_mz", "peak_intensity"):
        assert np.array_equal(getattr(sp, k), getattr(back, k)), k


def test_synthetic_generators_are_seeded():
    assert synth.synthetic_proteins(20, 3) == synth.synthetic_proteins(20, 3)
    a, _ = synth.synthetic_spectra(list(wl.proteins(30)), 5, 2, seed=11)
    b, _ = synth.synthetic_spectra(list(wl.proteins(30)), 5, 2, seed=11)
    assert np.array_equal(a.peak_mz, b.peak_mz) and np.array_equal(a.precursor_mz, b.precursor_mz)


def test_pack_proteins_and_subset():
    buf, off = maxdecoy.pack_proteins(["MKR", "", b"AAK"])
    assert buf.tobytes() == b"MKRAAK" and off.tolist() == [0, 3, 3, 6]
    sp, _ = wl.spectra(40, 6, 2)
    sub = sp.subset([4, 1])
    assert sub.spectrum_id.tolist() == [4, 1]
    assert np.array_equal(sub.peak_mz[:int(sub.peak_off[1])], sp.peak_mz[int(sp.peak_off[4]):int(sp.peak_off[5])])


@pytest.mark.parametrize("n,world,block", [(0, 2, 32), (1, 4, 32), (100, 2, 8), (1000, 8, 32), (37, 3, 5)])
def test_partition_is_a_balanced_permutation(n, world, block):
    rng = np.random.default_rng(n)
    mz, z = rng.uniform(300, 1500, n), rng.integers(2, 5, n)
    parts = parallel.partition_spectra(mz, z, world, block)
    allidx = np.concatenate(parts) if n else np.zeros(0, dtype=np.int64)
    assert sorted(allidx.tolist()) == list(range(n))
    assert max(len(p) for p in parts) <= parallel.padded_rows(n, world, block)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= block
    for p in parts:                                     # every shard ascends in neutral mass
        m = mz[p] * z[p] - 1.007276 * z[p]
        assert np.all(np.diff(m) >= 0)


def test_comm_entry_points_single_rank():
    """md_comm_init / md_gather_psms with one rank: no collective library needed, the gather is a copy, and the sharded
    driver over the library's own collective returns what a plain identify returns."""
    from oracle_lib import oracle_engine
    e = oracle_engine(2)
    e.digest(list(wl.proteins(60)), 2, 5, 50)
    e.set_modifications([synth.CAM], 0)
    e.index_build()
    sp, _ = wl.spectra(60, 20, 2)
    prm = maxdecoy.SearchParams(10, 10, n_decoys=10, seed=3, top_k=2)
    e.comm_init(0, 1, None)
    want, _ = e.identify(sp, prm)
    got, _ = parallel.identify_sharded_comm(e, sp, prm, 0, 1, block=8)
    assert got.tobytes() == want.tobytes()
    with pytest.raises(maxdecoy.MaxDecoyError):
        e.comm_init(1, 1, None)
    e.comm_destroy()
    e.close()


# ------------------------------------------------------------------------------------------ N > 1 (gloo, CPU)
WORKER = r"""
import os, sys, pickle
import numpy as np
sys.path.insert(0, os.path.join(%(root)r, "max-decoy_b200")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch.distributed as dist
from maxdecoy import SearchParams, parallel, synth
from oracle_lib import oracle_engine
import workloads as wl
dist.init_process_group("gloo")
e = oracle_engine(2)
e.digest(list(wl.proteins(120)), 2, 5, 50); e.set_modifications([synth.CAM, synth.OXM], 3); e.index_build()
sp, _ = wl.spectra(120, 50, 2, with_ox=True)
prm = SearchParams(10, 10, n_decoys=20, seed=5, top_k=3)
table, st = parallel.identify_sharded(e, sp, prm, block=8)
with open(os.path.join(%(out)r, "rank%%d.pkl" %% dist.get_rank()), "wb") as fh:
    pickle.dump((table, st["n_spectra"]), fh)
dist.destroy_process_group()
"""


def test_two_ranks_gather_equals_single_rank(tmp_path):
    """World size 2 over gloo: sharded identification + PSM all_gather gives, on every rank, exactly the table
    one rank computes alone (decoy RNG is keyed by the global spectrum id)."""
    import pickle
    from oracle_lib import oracle_engine
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": str(tmp_path)})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                           "--master-port", "29617", str(script)], env=env, timeout=300)
    e = oracle_engine(2)
    e.digest(list(wl.proteins(120)), 2, 5, 50)
    e.set_modifications([synth.CAM, synth.OXM], 3)
    e.index_build()
    sp, _ = wl.spectra(120, 50, 2, with_ox=True)
    single, _ = e.identify(sp, maxdecoy.SearchParams(10, 10, n_decoys=20, seed=5, top_k=3))
    got = []
    for r in range(2):
        with open(tmp_path / ("rank%d.pkl" % r), "rb") as fh:
            table, n_local = pickle.load(fh)
        assert n_local in (24, 26)                      # 50 spectra in blocks of 8 over 2 ranks
        got.append(table)
    assert got[0].tobytes() == got[1].tobytes()
    for f in single.dtype.names:
        assert np.array_equal(got[0][f], single[f]), f
