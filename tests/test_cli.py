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


def test_sequence_mass():
    out = subprocess.check_output(CLI + ["sequence-mass", "-s", "VVGTVK"], text=True)
    assert out.strip() == "601.379894"                   # tasks/sequence_mass.rs:24-27


def test_spectrum_splitup(tmp_path):
    sp, _ = wl.spectra(40, 4, 2)
    src = tmp_path / "run.mzML"
    mzml.write_mzml(sp, str(src))
    subprocess.check_call(CLI + ["spectrum-splitup", "-m", str(src), "-d", str(tmp_path / "split"), "-s", "_x"])
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
