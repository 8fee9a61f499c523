"""Host-side mirror of the reference's interface for the identification hot path.

Names follow the reference (src/proteomic): `Modification` (models/amino_acids/modification.rs),
`AminoAcid.get_sequence_weight` (models/amino_acids/amino_acid.rs:130), `mass.*`
(models/mass/mod.rs), `Engine.digest` (models/enzyms/digest_enzym.rs:32), `Engine.identify`
(tasks/identification.rs:160), `Engine.generate_decoys` (utility/decoy_generator.rs:108).
Everything computes through the C ABI of include/maxdecoy.h; there is no Python arithmetic on
the data path and no CPU fallback: `load()` raises if the CUDA library is not built.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import ALPHABET

_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB = os.path.join(os.path.dirname(_HERE), "csrc", "libmaxdecoy_cuda.so")

_lib = None


class MaxDecoyError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (_abi.STATUS_NAMES.get(status, status), message))
        self.status = status


def load(path=None):
    """Load the CUDA implementation (the product).  Fails loudly if it is not built."""
    global _lib
    if path is None:
        if _lib is not None:
            return _lib
        if not os.path.exists(CUDA_LIB):
            raise ImportError("CUDA extension missing: %s (run `python -c 'import __graft_entry__ as g; g.build()'`)" % CUDA_LIB)
        _lib = _abi.bind(CUDA_LIB)
        return _lib
    return _abi.bind(path)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _take(ptr, n, dtype):
    """Copy n elements of a library-owned array into numpy."""
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


class Modification:
    """One row of the modification CSV (modification.rs:36-76): accession,name,position,is_fix,aa,mono_mass."""

    def __init__(self, accession, name, position, is_fix, amino_acid, mono_mass):
        self.accession = accession.strip().lower()     # modification.rs:48
        self.name = name.strip()
        self.position = position.strip().upper()[:1]
        self.is_fix = bool(is_fix)
        self.amino_acid = amino_acid.strip().upper()[:1]
        self.mono_mass = float(mono_mass)

    @property
    def mono_mass_int(self):
        return mass.convert_mass_to_int(self.mono_mass)

    @classmethod
    def create_from_csv_file(cls, path):
        """modification.rs:115-130: csv::Reader with a header row, 6 columns."""
        import csv
        out = []
        with open(path, newline="") as fh:
            rows = list(csv.reader(fh))
        for row in rows[1:]:
            if not row:
                continue
            if len(row) != 6:
                raise ValueError("row has wrong length")   # modification.rs:59-61
            out.append(cls(row[0], row[1], row[2], int(row[3].strip()) > 0, row[4], float(row[5].strip())))
        return out

    def to_c(self):
        m = _abi.md_modification()
        m.accession = self.accession.encode()[:23]
        m.name = self.name.encode()[:39]
        m.position = ord(self.position)
        m.is_fix = 1 if self.is_fix else 0
        m.amino_acid = ord(self.amino_acid)
        m.mono_mass = self.mono_mass_int
        return m


class mass:
    """models/mass/mod.rs"""

    @staticmethod
    def convert_mass_to_int(m):
        return int(np.float64(m) * np.float64(1000000.0))      # truncation, mass/mod.rs:6-8

    @staticmethod
    def convert_mass_to_float(m):
        return m / 1000000.0


class Spectra:
    """Structure-of-arrays MS2 spectra (precursor m/z + charge as in utility/mz_ml/spectrum.rs:33-103, plus peaks)."""

    def __init__(self, precursor_mz, charge, peak_off, peak_mz, peak_intensity, spectrum_id=None):
        self.precursor_mz = np.ascontiguousarray(precursor_mz, dtype=np.float64)
        self.charge = np.ascontiguousarray(charge, dtype=np.uint8)
        self.peak_off = np.ascontiguousarray(peak_off, dtype=np.uint64)
        self.peak_mz = np.ascontiguousarray(peak_mz, dtype=np.float64)
        self.peak_intensity = np.ascontiguousarray(peak_intensity, dtype=np.float32)
        self.spectrum_id = None if spectrum_id is None else np.ascontiguousarray(spectrum_id, dtype=np.uint32)
        assert len(self.peak_off) == len(self.precursor_mz) + 1 == len(self.charge) + 1

    def __len__(self):
        return len(self.precursor_mz)

    def subset(self, idx):
        idx = np.asarray(idx)
        lens = (self.peak_off[1:] - self.peak_off[:-1])[idx].astype(np.int64)
        off = np.zeros(len(idx) + 1, dtype=np.uint64)
        off[1:] = np.cumsum(lens)
        sel = np.concatenate([np.arange(int(self.peak_off[i]), int(self.peak_off[i + 1])) for i in idx]) if len(idx) else np.zeros(0, dtype=np.int64)
        sid = self.spectrum_id[idx] if self.spectrum_id is not None else np.asarray(idx, dtype=np.uint32)
        return Spectra(self.precursor_mz[idx], self.charge[idx], off, self.peak_mz[sel], self.peak_intensity[sel], sid)

    def to_c(self):
        s = _abi.md_spectra()
        s.n = len(self)
        s.precursor_mz = self.precursor_mz.ctypes.data
        s.charge = self.charge.ctypes.data
        s.spectrum_id = self.spectrum_id.ctypes.data if self.spectrum_id is not None else None
        s.peak_off = self.peak_off.ctypes.data
        s.peak_mz = self.peak_mz.ctypes.data
        s.peak_intensity = self.peak_intensity.ctypes.data
        return s

    @property
    def nbytes(self):
        n = self.precursor_mz.nbytes + self.charge.nbytes + self.peak_off.nbytes + self.peak_mz.nbytes + self.peak_intensity.nbytes
        return n + (self.spectrum_id.nbytes if self.spectrum_id is not None else 0)


class SearchParams:
    """Flags of the `identification` subcommand (src/main.rs:385-496) that reach the hot path."""

    def __init__(self, lower_ppm=5, upper_ppm=5, fragment_tolerance=0.02, n_decoys=1000,
                 decoy_mode=_abi.DECOY_REFERENCE_RANDOM, seed=0, top_k=5, min_peaks=10,
                 max_fragment_charge=3, abs_lower_uda=0, abs_upper_uda=0, keep_decoys=False):
        self.lower_ppm, self.upper_ppm = int(lower_ppm), int(upper_ppm)
        self.fragment_tolerance = float(fragment_tolerance)
        self.n_decoys, self.decoy_mode, self.seed = int(n_decoys), int(decoy_mode), int(seed)
        self.top_k, self.min_peaks, self.max_fragment_charge = int(top_k), int(min_peaks), int(max_fragment_charge)
        self.abs_lower_uda, self.abs_upper_uda = int(abs_lower_uda), int(abs_upper_uda)
        self.keep_decoys = bool(keep_decoys)

    def to_c(self):
        p = _abi.md_search_params()
        p.lower_ppm, p.upper_ppm = self.lower_ppm, self.upper_ppm
        p.abs_lower_uda, p.abs_upper_uda = self.abs_lower_uda, self.abs_upper_uda
        p.fragment_tolerance = self.fragment_tolerance
        p.n_decoys, p.decoy_mode, p.seed = self.n_decoys, self.decoy_mode, self.seed
        p.top_k, p.min_peaks, p.max_fragment_charge = self.top_k, self.min_peaks, self.max_fragment_charge
        p.keep_decoys = 1 if self.keep_decoys else 0
        return p


def pack_proteins(sequences):
    """Concatenate protein sequences (bytes/str) into the residue buffer + offsets md_digest takes."""
    seqs = [s.encode() if isinstance(s, str) else bytes(s) for s in sequences]
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, dtype=np.uint8)
    return buf, off


class Engine:
    """One md_ctx (one device).  The call order mirrors the reference's workflow:
    `digest` (max_decoy digest) -> `set_modifications` + `index_build` -> `identify` (max_decoy identification)."""

    def __init__(self, lib=None, device=0, n_threads=0):
        self.lib = lib if lib is not None else load()
        cfg = _abi.md_config(device, n_threads)
        h = _abi.ctx_p()
        rc = self.lib.md_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise MaxDecoyError(rc, (self.lib.md_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.md_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise MaxDecoyError(rc, (self.lib.md_last_error(self.h) or b"").decode())

    @property
    def backend(self):
        return self.lib.md_backend_name().decode()

    # ---- masses (pure) -----------------------------------------------------------------
    def residue_mass(self, letter):
        return int(self.lib.md_residue_mass(ord(letter)))

    def get_sequence_weight(self, sequence):
        """AminoAcid::get_sequence_weight (amino_acid.rs:130-136)."""
        b = sequence.encode() if isinstance(sequence, str) else sequence
        return int(self.lib.md_sequence_weight(b, len(b)))

    def precursor_window(self, mz, charge, lower_ppm, upper_ppm):
        P, lo, hi = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.lib.md_precursor_window(mz, charge, lower_ppm, upper_ppm, C.byref(P), C.byref(lo), C.byref(hi)))
        return P.value, lo.value, hi.value

    # ---- modifications -----------------------------------------------------------------
    def set_modifications(self, mods, max_variable_mods=0):
        arr = (_abi.md_modification * max(1, len(mods)))(*[m.to_c() for m in mods])
        self._ck(self.lib.md_set_modifications(self.h, arr, len(mods), max_variable_mods))

    def substitution_map(self):
        """get_one_amino_acid_substitute_map (decoy_generator.rs:301-324) as a 21x21 array."""
        out = np.zeros(21 * 21, dtype=np.int64)
        self._ck(self.lib.md_substitution_map(self.h, out.ctypes.data_as(_abi.i64p)))
        return out.reshape(21, 21)

    # ---- digest ------------------------------------------------------------------------
    def digest(self, sequences, max_missed_cleavages=2, min_len=5, max_len=50):
        """DigestEnzym::digest for every protein (digest_enzym.rs:32-95); returns #unique peptides."""
        buf, off = pack_proteins(sequences)
        return self.digest_packed(buf, off, max_missed_cleavages, min_len, max_len)

    def digest_packed(self, residues, offsets, max_missed_cleavages=2, min_len=5, max_len=50):
        p = _abi.md_digest_params(max_missed_cleavages, min_len, max_len)
        n = C.c_uint64()
        self._ck(self.lib.md_digest(self.h, _ptr(residues), _ptr(offsets), len(offsets) - 1, C.byref(p), C.byref(n)))
        return n.value

    def peptides(self):
        t = _abi.md_peptide_table()
        self._ck(self.lib.md_peptides_export(self.h, C.byref(t)))
        try:
            n = t.n
            out = {
                "seq": _take(t.seq, t.seq_bytes, np.uint8),
                "seq_off": _take(t.seq_off, n + 1, np.uint64),
                "missed_cleavages": _take(t.missed_cleavages, n, np.uint8),
                "weight": _take(t.weight, n, np.int64),
                "counts": _take(t.counts, n * 21, np.int16).reshape(-1, 21),
                "assoc_off": _take(t.assoc_off, n + 1, np.uint64),
                "assoc_protein": _take(t.assoc_protein, t.n_assoc, np.uint32),
            }
        finally:
            self.lib.md_peptide_table_free(C.byref(t))
        return out

    @staticmethod
    def sequences_of(table):
        raw = table["seq"].tobytes()
        off = table["seq_off"]
        return [raw[int(off[i]):int(off[i + 1])].decode() for i in range(len(off) - 1)]

    # ---- index / lookup ----------------------------------------------------------------
    def index_build(self):
        self._ck(self.lib.md_index_build(self.h))

    def index_stats(self):
        s = _abi.md_index_stats()
        self._ck(self.lib.md_index_stats_get(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def window_search(self, lo, hi):
        lo = np.ascontiguousarray(lo, dtype=np.int64)
        hi = np.ascontiguousarray(hi, dtype=np.int64)
        b = np.zeros(len(lo), dtype=np.uint64)
        e = np.zeros(len(lo), dtype=np.uint64)
        self._ck(self.lib.md_window_search(self.h, _ptr(lo), _ptr(hi), len(lo), _ptr(b), _ptr(e)))
        return b, e

    def index_export(self, begin, count):
        pid = np.zeros(count, dtype=np.uint64)
        key = np.zeros(count, dtype=np.int64)
        self._ck(self.lib.md_index_export(self.h, begin, count, _ptr(pid), _ptr(key)))
        return pid, key

    @staticmethod
    def _precursors(precursors):
        arr = (_abi.md_precursor * max(1, len(precursors)))()
        for i, p in enumerate(precursors):
            arr[i].mass, arr[i].lo, arr[i].hi, arr[i].charge, arr[i].spectrum_id = [int(x) for x in p]
        return arr

    def candidates(self, precursors):
        """precursors: iterable of (P, lo, hi, charge, spectrum_id).  Returns CSR dict."""
        arr = self._precursors(precursors)
        t = _abi.md_candidate_table()
        self._ck(self.lib.md_candidates(self.h, arr, len(precursors), C.byref(t)))
        try:
            out = {"off": _take(t.off, t.n_spectra + 1, np.uint64), "peptide_id": _take(t.peptide_id, t.n, np.uint64),
                   "var_mask": _take(t.var_mask, t.n, np.uint64), "mod_weight": _take(t.mod_weight, t.n, np.int64)}
        finally:
            self.lib.md_candidate_table_free(C.byref(t))
        return out

    # ---- decoys ------------------------------------------------------------------------
    def _decoy_table(self, t):
        try:
            out = {"off": _take(t.off, t.n_spectra + 1, np.uint64), "seq": _take(t.seq, t.seq_bytes, np.uint8),
                   "seq_off": _take(t.seq_off, t.n + 1, np.uint64), "var_mask": _take(t.var_mask, t.n, np.uint64),
                   "weight": _take(t.weight, t.n, np.int64), "mod_weight": _take(t.mod_weight, t.n, np.int64),
                   "attempt": _take(t.attempt, t.n, np.uint32)}
        finally:
            self.lib.md_decoy_table_free(C.byref(t))
        return out

    def set_variable_mode(self, mode):
        """VARMOD_REFERENCE: the reference's all-or-first-hit placement (modified_peptide.rs:512-543);
        VARMOD_EXPANDED: every placement of up to nvar variable modifications is a candidate.  `index_build` afterwards."""
        self._ck(self.lib.md_set_variable_mode(self.h, int(mode)))

    def set_decoy_store(self, sequences):
        """The `decoys` table: decoys persisted by earlier runs, reused before new ones are generated
        (tasks/identification.rs:259-283; models/peptides/decoy.rs:118-153).  An empty list clears the store."""
        buf, off = pack_proteins(sequences)
        self._ck(self.lib.md_decoy_store_set(self.h, _ptr(buf), _ptr(off), len(off) - 1))

    def generate_decoys(self, precursors, n_per_spectrum, mode=_abi.DECOY_REFERENCE_RANDOM, seed=0):
        """DecoyGenerator::generate_decoys (decoy_generator.rs:108-219) for many spectra at once."""
        arr = self._precursors(precursors)
        t = _abi.md_decoy_table()
        self._ck(self.lib.md_generate_decoys(self.h, arr, len(precursors), n_per_spectrum, mode, seed, C.byref(t)))
        return self._decoy_table(t)

    def last_decoys(self):
        t = _abi.md_decoy_table()
        self._ck(self.lib.md_last_decoys_export(self.h, C.byref(t)))
        return self._decoy_table(t)

    # ---- identify ----------------------------------------------------------------------
    def identify(self, spectra, params, want_all_scores=False):
        """identification_task for a batch (identification.rs:201-368) incl. b/y scoring.
        Host buffers in, host PSM rows out.  Returns (psms structured array [n, top_k], stats[, scores, off])."""
        n, k = len(spectra), params.top_k
        psms = np.zeros((n, k), dtype=_abi.PSM_DTYPE)
        s, p = spectra.to_c(), params.to_c()
        st = _abi.md_identify_stats()
        if want_all_scores:
            sc, off = _abi.i64p(), _abi.u64p()
            self._ck(self.lib.md_identify(self.h, C.byref(s), C.byref(p), _ptr(psms), C.byref(st), C.byref(sc), C.byref(off)))
            try:
                offs = _take(off, n + 1, np.uint64)
                scores = _take(sc, int(offs[-1]) if n else 0, np.int64)
            finally:
                self.lib.md_free(sc)
                self.lib.md_free(off)
            return psms, self._stats(st), scores, offs
        self._ck(self.lib.md_identify(self.h, C.byref(s), C.byref(p), _ptr(psms), C.byref(st), None, None))
        return psms, self._stats(st)

    def identify_device(self, spectra_c, params, psms_dev_ptr):
        """Device-resident variant: `spectra_c` is an md_spectra of device pointers, `psms_dev_ptr` a device
        buffer of n*top_k md_psm rows (e.g. the NCCL send buffer).  Asynchronous until `sync()`."""
        st = _abi.md_identify_stats()
        p = params.to_c()
        self._ck(self.lib.md_identify_device(self.h, C.byref(spectra_c), C.byref(p), C.c_void_p(psms_dev_ptr), C.byref(st)))
        return self._stats(st)

    def sync(self):
        self._ck(self.lib.md_sync(self.h))

    # ---- multi-GPU (one Engine per rank; spectra sharded, index replicated) ---------------
    def comm_unique_id(self):
        """Rank 0: the communicator id (128 bytes) to hand to the other ranks."""
        buf = (C.c_uint8 * _abi.COMM_ID_BYTES)()
        rc = self.lib.md_comm_unique_id(buf)
        if rc != 0:
            raise MaxDecoyError(rc, (self.lib.md_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, rank, nranks, unique_id=None):
        buf = (C.c_uint8 * _abi.COMM_ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        self._ck(self.lib.md_comm_init(self.h, int(rank), int(nranks), buf))
        self.rank, self.nranks = int(rank), int(nranks)

    def gather_psms(self, psms):
        """All-gather of this rank's [rows, top_k] PSM table (host array) -> [nranks * rows, top_k], rank major."""
        nranks = getattr(self, "nranks", 1)
        psms = np.ascontiguousarray(psms)
        out = np.zeros((nranks * psms.shape[0],) + psms.shape[1:], dtype=psms.dtype)
        rows = psms.size
        self._ck(self.lib.md_gather_psms(self.h, _ptr(psms), rows, _ptr(out)))
        return out

    def gather_psms_device(self, local_dev_ptr, rows, all_dev_ptr):
        """Device-pointer variant: asynchronous on the ctx stream until `sync()`."""
        self._ck(self.lib.md_gather_psms(self.h, C.c_void_p(local_dev_ptr), int(rows), C.c_void_p(all_dev_ptr)))

    def comm_destroy(self):
        self._ck(self.lib.md_comm_destroy(self.h))

    @staticmethod
    def _stats(st):
        return {k: getattr(st, k) for k, _ in st._fields_}
