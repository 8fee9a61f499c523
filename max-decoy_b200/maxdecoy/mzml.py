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


def write_mzml(spectra, path=None, compress=True, indexed=False):
    """Minimal mzML writer for tests and for feeding the generated candidate sets to Comet.  `indexed` wraps the document in
    an <indexedmzML> root (without an index) -- the nesting `spectrum-splitup` expects of its input."""
    def enc(a, dtype):
        raw = np.asarray(a, dtype=dtype).tobytes()
        if compress:
            raw = zlib.compress(raw)
        return base64.b64encode(raw).decode()
    out = ['<?xml version="1.0" encoding="utf-8"?>']
    if indexed:
        out.append('<indexedmzML xmlns="http://psi.hupo.org/ms/mzml">')
    out += ['<mzML xmlns="http://psi.hupo.org/ms/mzml" version="1.1.0">', '<run id="synthetic"><spectrumList count="%d">' % len(spectra)]
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
    if indexed:
        out.append('</indexedmzML>')
    text = "\n".join(out) + "\n"
    if path:
        with open(path, "w") as fh:
            fh.write(text)
    return text


# ------------------------------------------------------------------------------------------------
# spectrum-splitup in the reference's own output format
# ------------------------------------------------------------------------------------------------
_INDENT = "    "                       # utility/mz_ml/mod.rs:6
_TOKEN = re.compile(r"<\?.*?\?>|<!--.*?-->|<!\[CDATA\[.*?\]\]>|<![^>]*>|</[^>]*>|<[^>]*?/>|<[^>]*>|[^<]+", re.S)
# url::percent_encoding::DEFAULT_ENCODE_SET: C0 controls, space, " # < > ` ? { } and everything above '~'
_KEEP = set(chr(c) for c in range(0x21, 0x7F)) - set('"#<>`?{}')


def _events(text):
    """(kind, raw) for every XML event quick-xml would report: 'decl', 'start', 'empty', 'end', 'text', 'other'."""
    for m in _TOKEN.finditer(text):
        t = m.group(0)
        if t.startswith("<?"):
            yield "decl", t
        elif t.startswith("<!"):
            yield "other", t
        elif t.startswith("</"):
            yield "end", t
        elif t.endswith("/>"):
            yield "empty", t
        elif t.startswith("<"):
            yield "start", t
        else:
            yield "text", t


def _name(tag):
    return re.match(r"</?\s*([^\s/>]+)", tag).group(1)


def _attr(tag, key):
    m = re.search(r'\b%s\s*=\s*"([^"]*)"' % re.escape(key), tag) or re.search(r"\b%s\s*=\s*'([^']*)'" % re.escape(key), tag)
    return m.group(1) if m else None


def content_before_spectrum_list(text):
    """MzMlReader::get_content_before_spectrum_list (mz_ml_reader.rs:102-143): every declaration, start, empty and end tag
    in front of <spectrumList>, one per line, indented by nesting level; text nodes are dropped."""
    out, level = [], 0
    for kind, raw in _events(text):
        if kind == "start":
            if _name(raw) == "spectrumList":
                break
            out.append(_INDENT * level + raw + "\n")
            level += 1
        elif kind in ("empty", "decl"):
            out.append(_INDENT * level + raw + "\n")
        elif kind == "end":
            level -= 1
            out.append(_INDENT * level + raw + "\n")
    return "".join(out)


def ms_two_spectrum_xml(text):
    """MzMlReader::get_ms_two_spectra (mz_ml_reader.rs:24-100): [(xml, indent_level)] of the spectra whose first `ms level`
    cvParam has value "2"; the xml is re-indented tag by tag, <binary> payloads stay on the line of their tag."""
    out, level, inside, in_binary, buf = [], 0, False, False, []
    for kind, raw in _events(text):
        if kind == "start":
            n = _name(raw)
            if n == "spectrum":
                inside = True
                buf.append(_INDENT * level + raw + "\n")
            elif n == "binary":
                buf.append(_INDENT * level + raw)
                in_binary = True
            elif inside:
                buf.append(_INDENT * level + raw + "\n")
            level += 1
        elif kind == "empty":
            if inside:
                buf.append(_INDENT * level + raw + "\n")
        elif kind == "text":
            if inside and in_binary:
                buf.append(raw)
        elif kind == "end":
            level -= 1
            n = _name(raw)
            if n == "spectrum" and inside:
                buf.append(_INDENT * level + raw)
                xml = "".join(buf)
                ms2 = None
                for k2, r2 in _events(xml):                       # is_ms_two_spectrum: the first cvParam named "ms level" decides
                    if k2 == "empty" and _name(r2) == "cvParam" and _attr(r2, "name") == "ms level" and _attr(r2, "value") is not None:
                        ms2 = _attr(r2, "value") == "2"
                        break
                if ms2:
                    out.append((xml, level))
                buf, inside = [], False
            elif n == "binary":
                in_binary = False
                buf.append(raw + "\n")
            elif inside:
                buf.append(_INDENT * level + raw + "\n")
    return out


def split_filename(spectrum_id, file_suffix=""):
    """Spectrum::get_filename + the suffix / extension handling of to_mz_ml (spectrum.rs:160-167,215-222): the scan id (or the
    whole id) percent-encoded, `_suffix`, and Path::set_extension("mzML") -- which REPLACES whatever follows the last dot."""
    name = scan_id_of_reference(spectrum_id) or spectrum_id
    enc = "".join(c if c in _KEEP else "".join("%%%02X" % b for b in c.encode()) for c in name)
    if file_suffix:
        enc += "_" + file_suffix
    stem = enc
    if "." in enc[1:]:
        stem = enc[:enc.rindex(".")]
    return stem + ".mzML"


def scan_id_of_reference(spectrum_id):
    """Spectrum::exctract_scan_id_from_spectrum_id (spectrum.rs:243-255): the match of `scan=.*\\b` (greedy, to the last word
    boundary), then what follows its last '='."""
    best = None
    i = spectrum_id.find("scan=")
    if i < 0:
        return ""
    tail = spectrum_id[i:]
    for end in range(len(tail), 4, -1):          # longest prefix of the tail that ends at a word boundary
        a = tail[end - 1]
        b = tail[end] if end < len(tail) else ""
        wa, wb = a.isalnum() or a == "_", (b.isalnum() or b == "_") if b else False
        if wa != wb:
            best = tail[:end]
            break
    return best.split("=")[-1] if best is not None else ""


def to_mz_ml(xml, indent_level, spectrum_id, before):
    """Spectrum::to_mz_ml (spectrum.rs:170-231): an indexedmzML file with this one spectrum -- byte offsets of the spectrum and
    of the index list, SHA-1 over everything up to and including the opening <fileChecksum> tag."""
    import hashlib
    index_list = _INDENT + '<indexList count="1">\n' + _INDENT * 2 + '<index name="spectrum">\n'
    content = before + _INDENT * (indent_level - 1) + '<spectrumList count="1" defaultDataProcessingRef="pwiz_Reader_conversion">\n'
    offset = len(content.encode()) + xml.index("<")
    index_list += _INDENT * 3 + '<offset idRef="%s">%d</offset>\n' % (spectrum_id, offset)
    content += xml + "\n" + _INDENT * (indent_level - 1) + "</spectrumList>\n" + _INDENT * 2 + "</run>\n" + _INDENT + "</mzML>\n"
    index_list += _INDENT * 2 + "</index>\n" + _INDENT + "</indexList>\n"
    offset = len(content.encode()) + index_list.index("<")
    content += index_list + _INDENT + "<indexListOffset>%d</indexListOffset>\n" % offset + _INDENT + "<fileChecksum>"
    digest = hashlib.sha1(content.encode()).hexdigest()
    return content + digest + "</fileChecksum>\n</indexedmzML>"


def spectrum_splitup(text, destination_folder, file_suffix=""):
    """`spectrum-splitup` (src/main.rs:183-206): one indexedmzML file per MS2 spectrum, named and laid out as the reference
    does it.  Returns the file names."""
    import os
    before = content_before_spectrum_list(text)
    os.makedirs(destination_folder, exist_ok=True)
    names = []
    for xml, level in ms_two_spectrum_xml(text):
        first = next(r for k, r in _events(xml) if k == "start")
        sid = _attr(first, "id")
        if sid is None:
            raise ValueError("spectrum without id")              # the reference panics (spectrum.rs:84-90)
        name = split_filename(sid, file_suffix)
        with open(os.path.join(destination_folder, name), "w", newline="") as fh:
            fh.write(to_mz_ml(xml, level, sid, before))
        names.append(name)
    return names
