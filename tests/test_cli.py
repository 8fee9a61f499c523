"""`max_decoy` command line (the reference's subcommands / flags, src/main.rs:213-496)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from maxdecoy import mzml, synth
import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = [sys.executable, os.path.join(ROOT, "max-decoy_b200", "max_decoy.py")]


HOST = os.path.join(ROOT, "max-decoy_b200", "host", "max_decoy")


def _native_host():
    if not os.path.exists(HOST):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "max-decoy_b200", "csrc"), "-j8", "-s"])
        subprocess.check_call(["make", "-C", os.path.dirname(HOST), "-s"])
    return [HOST]


def test_sequence_mass():
    for cli in (CLI, _native_host()):
        out = subprocess.check_output(cli + ["sequence-mass", "-s", "VVGTVK"], text=True)
        assert out.strip() == "601.379894"               # tasks/sequence_mass.rs:24-27


def test_native_host_reports_errors_instead_of_falling_back(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    (tmp_path / "db.fasta").write_text(">sp|P00001|X\nMKAAAR\n")
    r = subprocess.run(_native_host() + ["digest", "-i", str(tmp_path / "db.fasta"), "-o", str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_spectrum_splitup(tmp_path):
    sp, _ = wl.spectra(40, 4, 2)
    src = tmp_path / "run.mzML"
    mzml.write_mzml(sp, str(src), indexed=True)
    subprocess.check_call(CLI + ["spectrum-splitup", "-m", str(src), "-d", str(tmp_path / "split"), "-s", "x"])
    files = sorted(os.listdir(tmp_path / "split"))
    assert files == ["1_x.mzML", "2_x.mzML", "3_x.mzML", "4_x.mzML"]
    one, ids = mzml.read_ms_two_spectra(str(tmp_path / "split" / "3_x.mzML"))
    assert len(one) == 1 and one.precursor_mz[0] == sp.precursor_mz[2]
    assert np.array_equal(one.peak_mz, sp.peak_mz[int(sp.peak_off[2]):int(sp.peak_off[3])])


@pytest.mark.gpu
def test_digest_and_identification_end_to_end(tmp_path):
    prots = list(wl.proteins(150))
    (tmp_path / "db.fasta").write_text(synth.fasta_text(prots))
    (tmp_path / "mods.csv").write_text(synth.mods_csv_text([synth.CAM, synth.OXM]))
    sp, truth = wl.spectra(150, 12, 2, with_ox=True)
    (tmp_path / "run.mgf").write_text(synth.mgf_text(sp))
    subprocess.check_call(CLI + ["digest", "-i", str(tmp_path / "db.fasta"), "-c", "2", "-l", "5", "-h", "50", "-o", str(tmp_path / "dig")])
    assert len((tmp_path / "dig" / "peptides.csv").read_text().splitlines()) > 1000
    subprocess.check_call(CLI + ["identification", "-m", str(tmp_path / "mods.csv"), "-s", str(tmp_path / "run.mgf"), "--fasta", str(tmp_path / "db.fasta"),
                                 "-n", "3", "-d", "40", "-l", "10", "-u", "10", "--seed", "3", "-o", str(tmp_path / "out")])
    rows = (tmp_path / "out" / "psms.csv").read_text().splitlines()
    best = {r.split(",")[0]: r.split(",") for r in rows if r.split(",")[2] == "1"}
    hits = sum(1 for i, (seq, _) in enumerate(truth) if best.get("scan=%d" % (i + 1), [""] * 6)[5] == seq)
    assert hits >= 6                                       # most database spectra are identified by their generating peptide
    assert (tmp_path / "out" / "1.fasta").exists() and (tmp_path / "out" / "1.comet.params").exists()
    # the native (C++) host writes the same files from the same inputs
    subprocess.check_call(_native_host() + ["identification", "-m", str(tmp_path / "mods.csv"), "-s", str(tmp_path / "run.mgf"), "--fasta", str(tmp_path / "db.fasta"),
                                            "-n", "3", "-d", "40", "-l", "10", "-u", "10", "--seed", "3", "-o", str(tmp_path / "out_native")])
    for name in ("1.fasta", "7.fasta", "12.fasta"):
        assert (tmp_path / "out_native" / name).read_text() == (tmp_path / "out" / name).read_text(), name
    a = (tmp_path / "out_native" / "1.comet.params").read_text().replace("out_native", "out")
    assert a == (tmp_path / "out" / "1.comet.params").read_text()
    rows_native = (tmp_path / "out_native" / "psms.csv").read_text().splitlines()
    assert [r.split(",")[:10] for r in rows_native] == [r.split(",")[:10] for r in rows]
    # a later run finds the first run's decoys in its store (the `decoys` table) and reuses them before generating new ones
    # (tasks/identification.rs:259-283); both hosts agree again
    stored = {ln.split(",")[1] for ln in (tmp_path / "out" / "decoys.csv").read_text().splitlines()}
    assert len(stored) > 300
    common = ["identification", "-m", str(tmp_path / "mods.csv"), "-s", str(tmp_path / "run.mgf"), "--fasta", str(tmp_path / "db.fasta"),
              "-n", "3", "-d", "40", "-l", "10", "-u", "10", "--seed", "3", "--stored-decoys", str(tmp_path / "out" / "decoys.csv")]
    subprocess.check_call(CLI + common + ["-o", str(tmp_path / "out2")])
    subprocess.check_call(_native_host() + common + ["-o", str(tmp_path / "out2_native")])
    for name in ("1.fasta", "7.fasta", "12.fasta"):
        text = (tmp_path / "out2" / name).read_text()
        assert (tmp_path / "out2_native" / name).read_text() == text, name
        decoys = [ln for ln, prev in zip(text.splitlines()[1:], text.splitlines()) if prev.startswith(">DECOY_")]
        assert len(decoys) == 40 and set(decoys) <= stored, name
    subprocess.check_call(_native_host() + ["digest", "-i", str(tmp_path / "db.fasta"), "-c", "2", "-l", "5", "-h", "50", "-o", str(tmp_path / "dig_native")])
    assert (tmp_path / "dig_native" / "peptides.csv").read_text() == (tmp_path / "dig" / "peptides.csv").read_text()
    assert (tmp_path / "dig_native" / "peptides_proteins.csv").read_text() == (tmp_path / "dig" / "peptides_proteins.csv").read_text()


@pytest.mark.gpu
def test_identification_with_terminal_modifications_both_hosts(tmp_path):
    """Terminal (N / C) modifications through both command lines: same PSM rows, same per-spectrum FASTA (incl. the ModRes
    summary with the terminus slots, modified_peptide.rs:606-659) and Comet parameters."""
    from maxdecoy import Modification
    prots = list(wl.proteins(150))
    mods = [synth.CAM, synth.OXM, Modification("unimod:1", "Acetyl", "N", True, "M", 42.010565),
            Modification("unimod:28", "Gln->pyro-Glu", "N", False, "Q", -17.026549), Modification("x:259", "Label", "C", True, "K", 8.014199)]
    (tmp_path / "db.fasta").write_text(synth.fasta_text(prots))
    (tmp_path / "mods.csv").write_text(synth.mods_csv_text(mods))
    sp, _ = wl.spectra(150, 12, 2)
    (tmp_path / "run.mgf").write_text(synth.mgf_text(sp))
    common = ["identification", "-m", str(tmp_path / "mods.csv"), "-s", str(tmp_path / "run.mgf"), "--fasta", str(tmp_path / "db.fasta"),
              "-n", "2", "-d", "30", "-l", "2000", "-u", "2000", "--seed", "3"]
    subprocess.check_call(CLI + common + ["-o", str(tmp_path / "out")])
    subprocess.check_call(_native_host() + common + ["-o", str(tmp_path / "out_native")])
    rows = (tmp_path / "out" / "psms.csv").read_text().splitlines()
    rows_native = (tmp_path / "out_native" / "psms.csv").read_text().splitlines()
    assert len(rows) > 12 and [r.split(",")[:10] for r in rows_native] == [r.split(",")[:10] for r in rows]
    summaries = set()
    for name in ("1.fasta", "5.fasta", "12.fasta"):
        text = (tmp_path / "out" / name).read_text()
        assert (tmp_path / "out_native" / name).read_text() == text, name
        summaries |= {ln.split("ModRes=")[1] for ln in text.splitlines() if "ModRes=" in ln}
    assert any("x:259|Label" in s for s in summaries)          # a C-terminal K somewhere among targets / decoys
    a = (tmp_path / "out_native" / "1.comet.params").read_text().replace("out_native", "out")
    assert a == (tmp_path / "out" / "1.comet.params").read_text()
