"""The callers and data formats either side of the hot path: per-spectrum drop-in outputs,
mzML input, PostgreSQL CSV export.  CPU only; where an engine is needed the oracle stands in for the CUDA library
(same ABI), so these tests check host logic and output bytes, not the product's arithmetic."""
import json
import os

import numpy as np
import pytest

import maxdecoy
from maxdecoy import SearchParams, mzml, outputs, pgexport, synth
from oracle_lib import oracle_engine
import workloads as wl

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cpu():
    e = oracle_engine(4)
    yield e
    e.close()


# ------------------------------------------------------------------------------------------ headers / ModRes
def test_modification_summary_and_headers():
    mods = [synth.CAM, synth.OXM]
    # modified_peptide.rs:606-659: "(count|accession|name)", keys "accession|name" ascending; accession lower-cased
    assert outputs.modification_summary("ACMCK", mods, 0) == "(2|unimod:4|Carbamidomethyl)"
    assert outputs.modification_summary("ACMCMK", mods, 0b010100) == "(2|unimod:35|Oxidation)(2|unimod:4|Carbamidomethyl)"
    assert outputs.modification_summary("AAAK", mods, 0) == ""
    assert outputs.peptide_header("ACK", 7, "") == ">PEPTIDE_ACK MaxDecoyId=7"                       # peptide.rs:238-240
    assert outputs.peptide_header("ACK", 7, "(1|unimod:4|Carbamidomethyl)") == ">PEPTIDE_ACK MaxDecoyId=7 ModRes=(1|unimod:4|Carbamidomethyl)"
    assert outputs.decoy_header("WCK", "") == ">DECOY_WCK"                                           # decoy.rs:76-82
    assert outputs.decoy_header("WCK", "(1|unimod:4|Carbamidomethyl)") == ">DECOY_WCK ModRes=(1|unimod:4|Carbamidomethyl)"
    assert outputs.fasta_entry(">X", "AAK") == ">X\nAAK\n"                                           # fasta_entry.rs:16-18
    text, nt, nd = outputs.spectrum_fasta([("ACK", 1, 0), ("ACK", 9, 0), ("MMK", 2, 0b11)], [("WCK", 0), ("WCK", 0)], mods)
    assert (nt, nd) == (2, 1) and text.count(">") == 3                                               # dedup by sequence
    assert ">PEPTIDE_MMK MaxDecoyId=2 ModRes=(2|unimod:35|Oxidation)\nMMK\n" in text


def test_rust_float_display():
    assert outputs.rust_f64(0.02) == "0.02" and outputs.rust_f64(10.0) == "10" and outputs.rust_f64(57.021464) == "57.021464"
    assert outputs.rust_f64(1.0005) == "1.0005" and outputs.rust_f64(113.08406) == "113.08406"


def test_comet_params_bytes_match_reference_template():
    """comet_parameter::new (utility/comet_parameter.rs:96-124) around the reference's two constant blocks
    (fixture extracted from the reference file by tests/golden/make_comet_template.py)."""
    with open(os.path.join(GOLDEN, "comet_params_template.json")) as fh:
        t = json.load(fh)
    mods = [synth.CAM, synth.OXM]
    got = outputs.comet_params("# comet_version 2019.01 rev. 4", mods, "/tmp/scan_17.fasta", 1042, 3, 0.02, 5, 10)
    want = ("# comet_version 2019.01 rev. 4\n" + t["begin"] +
            "peptide_mass_tolerance = 10.0000\nfragment_bin_tol = 0.02\nnum_results = 1042\nnum_output_lines = 1042\n"
            "database_name = /tmp/scan_17.fasta\nadd_C_cysteine = 57.021464\nadd_J_user_amino_acid = 113.08406\n"
            "variable_mod01 = 15.994915 M 0 3 -1 0 0\n" + t["end"])
    assert got == want
    j = maxdecoy.Modification("x:1", "Jmod", "A", True, "J", 1.5)
    got = outputs.comet_params("rev", [j], "a.fasta", 3, 0, 1.0005, 20, 7)
    assert "add_J_user_amino_acid = 114.58406\n" in got and got.count("add_J_user_amino_acid") == 1
    assert "peptide_mass_tolerance = 20.0000\nfragment_bin_tol = 1.0005\n" in got


def test_write_identification_outputs(cpu, tmp_path):
    mods = [synth.CAM, synth.OXM]
    cpu.digest(list(wl.proteins(120)), 2, 5, 50)
    cpu.set_modifications(mods, 3)
    cpu.index_build()
    sp, _ = wl.spectra(120, 6, 2, with_ox=True)
    prm = SearchParams(10, 10, n_decoys=25, seed=4, top_k=3)
    names = ["scan_%d" % (i + 1) for i in range(len(sp))]
    psms, st = outputs.write_identification_outputs(str(tmp_path), names, cpu, sp, prm, mods, 3, "# comet_version 2019.01 rev. 4")
    ref_psms, _ = cpu.identify(sp, prm)
    assert psms.tobytes() == ref_psms.tobytes()
    pre = wl.precursors_of(cpu, sp)
    cand = cpu.candidates(pre)
    targets = set(cpu.sequences_of(cpu.peptides()))
    for s, name in enumerate(names):
        fasta = (tmp_path / (name + ".fasta")).read_text().splitlines()
        heads, seqs = fasta[0::2], fasta[1::2]
        nt = int(cand["off"][s + 1] - cand["off"][s])
        assert sum(h.startswith(">PEPTIDE_") for h in heads) == nt and all(h.startswith(">PEPTIDE_") for h in heads[:nt])
        dec = [q for h, q in zip(heads, seqs) if h.startswith(">DECOY_")]
        assert len(dec) == int(psms["n_decoys"][s, 0]) and not (set(dec) & targets)
        for h, q in zip(heads, seqs):
            assert h.split()[0] in (">PEPTIDE_" + q, ">DECOY_" + q)
            if "C" in q:
                assert "|unimod:4|Carbamidomethyl)" in h
        params = (tmp_path / (name + ".comet.params")).read_text()
        assert "num_results = %d\n" % len(heads) in params and "database_name = %s\n" % (tmp_path / (name + ".fasta")) in params
        assert (tmp_path / (name + ".less_decoys")).exists() == (len(dec) < 25)


# ------------------------------------------------------------------------------------------ mzML
@pytest.mark.parametrize("compress", [True, False])
def test_mzml_round_trip(compress):
    sp, _ = wl.spectra(40, 5, 2)
    text = mzml.write_mzml(sp, compress=compress)
    # an MS1 spectrum in between must be skipped (MzMlReader::is_ms_two_spectrum)
    text = text.replace('<spectrumList count="5">', '<spectrumList count="6"><spectrum index="99" id="scan=99" defaultArrayLength="0">'
                        '<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="1"/></spectrum>')
    back, ids = mzml.read_ms_two_spectra(text)
    assert len(back) == 5 and ids[0] == ("controllerType=0 controllerNumber=1 scan=1", "1")
    for k in ("precursor_mz", "charge", "peak_off", "peak_mz", "peak_intensity"):
        assert np.array_equal(getattr(sp, k), getattr(back, k)), k


def test_mzml_missing_precursor_is_an_error():
    bad = ('<mzML><run><spectrumList><spectrum id="scan=1"><cvParam name="ms level" value="2"/>'
           '<precursorList><precursor><selectedIonList><selectedIon><cvParam name="selected ion m/z" value="500.5"/>'
           '</selectedIon></selectedIonList></precursor></precursorList></spectrum></spectrumList></run></mzML>')
    with pytest.raises(ValueError):                      # spectrum.rs:84-90 panics
        mzml.read_ms_two_spectra(bad)


_SPLIT_SRC = """<?xml version="1.0" encoding="utf-8"?>
<indexedmzML xmlns="http://psi.hupo.org/ms/mzml">
  <mzML id="x" version="1.1.0">
    <cvList count="1">
      <cv id="MS" fullName="PSI MS"/>
    </cvList>
    <run id="r1">
      <spectrumList count="3" defaultDataProcessingRef="pwiz_Reader_conversion">
        <spectrum index="0" id="controllerType=0 controllerNumber=1 scan=7" defaultArrayLength="2">
          <cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="2"/>
          <precursorList count="1"><precursor><selectedIonList count="1"><selectedIon>
            <cvParam cvRef="MS" name="selected ion m/z" value="500.25"/>
            <cvParam cvRef="MS" name="charge state" value="2"/>
          </selectedIon></selectedIonList></precursor></precursorList>
          <binaryDataArrayList count="1"><binaryDataArray encodedLength="4"><binary>AAAA</binary></binaryDataArray></binaryDataArrayList>
        </spectrum>
        <spectrum index="1" id="scan=8" defaultArrayLength="0">
          <cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="1"/>
        </spectrum>
        <spectrum index="2" id="sample 9.raw" defaultArrayLength="0">
          <cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="2"/>
          <precursorList count="1"><precursor><selectedIonList count="1"><selectedIon>
            <cvParam cvRef="MS" name="selected ion m/z" value="600.5"/><cvParam cvRef="MS" name="charge state" value="3"/>
          </selectedIon></selectedIonList></precursor></precursorList>
        </spectrum>
      </spectrumList>
    </run>
  </mzML>
</indexedmzML>
"""


def test_spectrum_splitup_writes_the_reference_layout(tmp_path):
    """`spectrum-splitup` (src/main.rs:183-206; Spectrum::to_mz_ml, utility/mz_ml/spectrum.rs:170-231): one indexedmzML per
    MS2 spectrum -- header re-indented tag by tag (mz_ml_reader.rs:102-143), byte offsets of the spectrum and of the index
    list, SHA-1 over everything up to and including the opening <fileChecksum> tag, percent-encoded file names."""
    import hashlib
    import re
    names = mzml.spectrum_splitup(_SPLIT_SRC, str(tmp_path), "part")
    assert names == ["7_part.mzML", "sample%209.mzML"]       # scan id, or the encoded id; set_extension replaces ".raw_part"
    raw = (tmp_path / names[0]).read_bytes()
    text = raw.decode()
    assert text.startswith('<?xml version="1.0" encoding="utf-8"?>\n<indexedmzML xmlns="http://psi.hupo.org/ms/mzml">\n    <mzML id="x" version="1.1.0">\n'
                           '        <cvList count="1">\n            <cv id="MS" fullName="PSI MS"/>\n        </cvList>\n        <run id="r1">\n'
                           '            <spectrumList count="1" defaultDataProcessingRef="pwiz_Reader_conversion">\n'
                           '                <spectrum index="0" id="controllerType=0 controllerNumber=1 scan=7" defaultArrayLength="2">\n')
    assert '                            <binary>AAAA</binary>\n' in text          # payload stays on the line of its tag
    assert text.endswith('</fileChecksum>\n</indexedmzML>')
    off = int(re.search(r'<offset idRef="controllerType=0 controllerNumber=1 scan=7">(\d+)</offset>', text).group(1))
    assert raw[off:off + 9] == b"<spectrum"
    ioff = int(re.search(r"<indexListOffset>(\d+)</indexListOffset>", text).group(1))
    assert raw[ioff:ioff + 10] == b"<indexList"
    upto = raw.index(b"<fileChecksum>") + len(b"<fileChecksum>")
    assert hashlib.sha1(raw[:upto]).hexdigest().encode() == raw[upto:upto + 40]
    # the split files are what our own reader takes
    back, ids = mzml.read_ms_two_spectra(str(tmp_path / names[1]))
    assert len(back) == 1 and ids[0][0] == "sample 9.raw" and float(back.precursor_mz[0]) == 600.5 and int(back.charge[0]) == 3
    assert mzml.scan_id_of_reference("controllerType=0 controllerNumber=1 scan=7") == "7" and mzml.scan_id_of_reference("no scan here") == ""


# ------------------------------------------------------------------------------------------ PostgreSQL CSV
def test_pg_csv_exports(cpu):
    prots = list(wl.proteins(30))
    headers, seqs = synth.read_fasta(synth.fasta_text(prots))
    assert pgexport.extract_accession(headers[3]) == "P00003"          # protein.rs:27
    assert pgexport.extract_accession(">sp|Q9Y6K9|NEMO_HUMAN x") == "Q9Y6K9" and pgexport.extract_accession(">nothing") == ""
    rows = pgexport.proteins_csv(headers, seqs).splitlines()
    assert rows[0].split(",")[:2] == ["1", "P00000"] and rows[0].endswith(",t") and len(rows) == 30
    cpu.digest(prots, 2, 5, 50)
    t = cpu.peptides()
    pcsv = pgexport.peptides_csv(t).splitlines()
    assert len(pcsv) == len(t["weight"])
    k = len(pcsv) // 2
    f = pcsv[k].split(",")
    s = cpu.sequences_of(t)[k]
    assert len(f) == 26 and f[0] == str(k + 1) and f[1] == s and int(f[2]) == len(s) and int(f[4]) == int(t["weight"][k])
    # schema.sql:20-40: r n d c e q g h j k m f p o s t u v w y, a last
    assert [int(x) for x in f[5:]] == [s.count(c.upper()) for c in "rndceqghjkmfpostuvwya"]
    assoc = pgexport.peptides_proteins_csv(t).splitlines()
    assert len(assoc) == len(t["assoc_protein"]) and assoc[0] == "1,%d" % (int(t["assoc_protein"][0]) + 1)
    cpu.set_modifications([synth.CAM], 0)
    cpu.index_build()
    sp, _ = wl.spectra(30, 4, 2)
    d = cpu.generate_decoys(wl.precursors_of(cpu, sp), 10, maxdecoy.DECOY_REFERENCE_RANDOM, seed=1)
    dcsv = pgexport.decoys_csv(d).splitlines()
    assert len(dcsv) == len(set(wl.decoy_strings(d)))
    f = dcsv[0].split(",")
    assert f[3] == "0" and int(f[4]) == int(d["weight"][0]) and len(f) == 26
    assert "CREATE TABLE psms" in pgexport.PSMS_DDL and "PRIMARY KEY (spectrum_id, rank)" in pgexport.PSMS_DDL
    psms, _ = cpu.identify(sp, SearchParams(10, 10, n_decoys=10, seed=1, top_k=2))
    ids = [("scan=%d" % (i + 1), str(i + 1)) for i in range(len(sp))]
    out = pgexport.psms_csv(psms, ids, lambda s_, r: "SEQ", lambda s_, r: "", [p[0] for p in wl.precursors_of(cpu, sp)])
    lines = out.splitlines()
    assert len(lines) == int((psms["rank"] > 0).sum()) and lines[0].startswith("scan=1,1,1,")


def test_env_file_and_database_scripts(tmp_path):
    """`.env` -> PGSQL_URL as DatabaseConnection::get_database_url reads it (utility/database_connection.rs:8-20), the psql
    load script for the exported tables in dependency order, and the way back (decoys table -> stored decoys)."""
    env = tmp_path / ".env"
    env.write_text("# comment\nPGSQL_URL=postgres://u:p@localhost:5432/maxdecoy\nPGSQL_TLS_MODE=NONE\n")
    assert pgexport.database_url(str(env), environ={}) == "postgres://u:p@localhost:5432/maxdecoy"
    assert pgexport.database_url(str(env), environ={"PGSQL_URL": "postgres://other"}) == "postgres://other"      # the environment wins (dotenv)
    env.write_text('export PGSQL_URL="postgres://quoted/db"\n')
    assert pgexport.database_url(str(env), environ={}) == "postgres://quoted/db"
    env.write_text("PGSQL_TLS_MODE=NONE\n")
    with pytest.raises(RuntimeError, match="PGSQL_URL"):
        pgexport.database_url(str(env), environ={})
    with pytest.raises(RuntimeError, match="Could not load"):
        pgexport.database_url(str(tmp_path / "missing.env"), environ={})
    env.write_text("PGSQL_URL=postgres://u@h/db\n")
    script, cmd = pgexport.load_script(str(tmp_path), env_path=str(env))
    order = [ln.split()[1] for ln in script.splitlines() if ln.startswith("\\copy")]
    assert order == ["proteins", "peptides", "peptides_proteins", "decoys", "psms"]
    assert "a_count) FROM" in script and script.index("r_count") < script.index("a_count")            # schema order: a_count last
    assert cmd[:2] == ["psql", "postgres://u@h/db"] and "CREATE TABLE IF NOT EXISTS psms" in script
    assert "\\copy (SELECT id, aa_sequence" in pgexport.unload_script(str(tmp_path))
    assert pgexport.read_sequences_csv("1,AAGK,4,0,1,0\n2,DECJY,5,0,1\nbad\n") == ["AAGK", "DECJY"]
