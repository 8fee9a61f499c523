"""mzML input.

`read_ms_two_spectra` mirrors MzMlReader::get_ms_two_spectra + Spectrum::new
(utility/mz_ml/mz_ml_reader.rs:24-100, utility/mz_ml/spectrum.rs:33-103): keep spectra whose `ms level` cvParam is 2,
take `selected ion m/z` and `charge state` from the cvParams inside <selectedIon>, the scan id from the spectrum id.
The reference never decodes peaks (it hands the mzML to Comet); scoring on the GPU needs them, so the two
<binaryDataArray>s (m/z array MS:1000514, intensity array MS:1000515; 32/64-bit float, optional zlib) are decoded too.
"""
import base64
import re
import struct
import xml.etree.ElementTree as ET
import zlib

import numpy as np

from .api import Spectra

_NS = re.compile(r"^\{.*\}")


def _tag(el):
    return _NS.sub("", el.tag)


def scan_id_of(spectrum_id):
    """Spectrum::exctract_scan_id_from_spectrum_id: the value of `scan=` in the native id, or ''."""
    m = re.search(r"scan=(\d+)", spectrum_id)
    return m.group(1) if m else ""


def _decode_array(bda):
    bits, compressed, kind, payload = 64, False, None, ""
    for el in bda.iter():
        t = _tag(el)
        if t == "cvParam":
            acc, name = el.get("accession", ""), el.get("name", "")
            if acc == "MS:1000521" or name == "32-bit float":
                bits = 32
            elif acc == "MS:1000523" or name == "64-bit float":
                bits = 64
            elif acc == "MS:1000574" or name == "zlib compression":
                compressed = True
            elif acc == "MS:1000514" or name == "m/z array":
                kind = "mz"
            elif acc == "MS:1000515" or name == "intensity array":
                kind = "intensity"
        elif t == "binary":
            payload = (el.text or "").strip()
    raw = base64.b64decode(payload) if payload else b""
    if compressed and raw:
        raw = zlib.decompress(raw)
    arr = np.frombuffer(raw, dtype="<f4" if bits == 32 else "<f8").astype(np.float64)
    return kind, arr


def read_ms_two_spectra(path_or_text):
    """-> (Spectra, ids) with ids = list of (spectrum_id, scan_id).  Raises ValueError where the reference panics
    (MS2 spectrum without selected ion m/z, charge state or id: spectrum.rs:84-90)."""
    text = path_or_text
    if "<" not in path_or_text[:200]:
        with open(path_or_text) as fh:
            text = fh.read()
    root = ET.fromstring(text)
    pmz, charge, off, mzs, ints, ids = [], [], [0], [], [], []
    for sp in root.iter():
        if _tag(sp) != "spectrum":
            continue
        level, mz, z = None, None, None
        arrays = {}
        for el in sp:
            t = _tag(el)
            if t == "cvParam" and el.get("name") == "ms level":
                level = int(el.get("value"))
        if level != 2:                                         # MzMlReader::is_ms_two_spectrum
            continue
        for sel in sp.iter():
            if _tag(sel) == "selectedIon":
                for cv in sel.iter():
                    if _tag(cv) != "cvParam":
                        continue
                    if cv.get("name") == "selected ion m/z":
                        mz = float(cv.get("value"))
                    elif cv.get("name") == "charge state":
                        z = int(cv.get("value"))
            elif _tag(sel) == "binaryDataArray":
                kind, arr = _decode_array(sel)
                if kind:
                    arrays[kind] = arr
        sid = sp.get("id")
        if mz is None or z is None or sid is None:
            raise ValueError("MS2 spectrum without selected ion m/z, charge state or id: %r" % sid)
        if not (0 < z < 256):
            raise ValueError("charge state out of u8 range in %r" % sid)
        m, i = arrays.get("mz", np.zeros(0)), arrays.get("intensity", np.zeros(0))
        n = min(len(m), len(i))
        order = np.argsort(m[:n], kind="stable")
        pmz.append(mz)
        charge.append(z)
        mzs.append(m[:n][order])
        ints.append(i[:n][order])
        off.append(off[-1] + n)
        ids.append((sid, scan_id_of(sid)))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt)
    return Spectra(np.array(pmz, dtype=np.float64), np.array(charge, dtype=np.uint8), np.array(off, dtype=np.uint64),
                   cat(mzs, np.float64), cat(ints, np.float32)), ids


def write_mzml(spectra, path=None, compress=True):
    """Minimal indexed-less mzML writer for tests and for feeding the generated candidate sets to Comet."""
    def enc(a, dtype):
        raw = np.asarray(a, dtype=dtype).tobytes()
        if compress:
            raw = zlib.compress(raw)
        return base64.b64encode(raw).decode()
    out = ['<?xml version="1.0" encoding="utf-8"?>', '<mzML xmlns="http://psi.hupo.org/ms/mzml" version="1.1.0">',
           '<run id="synthetic"><spectrumList count="%d">' % len(spectra)]
    for s in range(len(spectra)):
        a, b = int(spectra.peak_off[s]), int(spectra.peak_off[s + 1])
        out.append('<spectrum index="%d" id="controllerType=0 controllerNumber=1 scan=%d" defaultArrayLength="%d">' % (s, s + 1, b - a))
        out.append('<cvParam cvRef="MS" accession="MS:1000511" name="ms level" value="2"/>')
        out.append('<precursorList count="1"><precursor><selectedIonList count="1"><selectedIon>')
        out.append('<cvParam cvRef="MS" accession="MS:1000744" name="selected ion m/z" value="%r"/>' % float(spectra.precursor_mz[s]))
        out.append('<cvParam cvRef="MS" accession="MS:1000041" name="charge state" value="%d"/>' % int(spectra.charge[s]))
        out.append('</selectedIon></selectedIonList></precursor></precursorList>')
        out.append('<binaryDataArrayList count="2">')
        for acc, name, data, dt, bits in (("MS:1000514", "m/z array", spectra.peak_mz[a:b], "<f8", ("MS:1000523", "64-bit float")),
                                          ("MS:1000515", "intensity array", spectra.peak_intensity[a:b], "<f4", ("MS:1000521", "32-bit float"))):
            out.append('<binaryDataArray><cvParam cvRef="MS" accession="%s" name="%s"/>' % bits)
            if compress:
                out.append('<cvParam cvRef="MS" accession="MS:1000574" name="zlib compression"/>')
            out.append('<cvParam cvRef="MS" accession="%s" name="%s"/><binary>%s</binary></binaryDataArray>' % (acc, name, enc(data, dt)))
        out.append('</binaryDataArrayList></spectrum>')
    out.append('</spectrumList></run></mzML>')
    text = "\n".join(out) + "\n"
    if path:
        with open(path, "w") as fh:
            fh.write(text)
    return text
