"""Pure-Python second restatement of the reference's arithmetic, used ONLY to pin the C++ oracle on small cases.

Every function follows the reference text literally (file:line under /root/reference/src/proteomic) and is
deliberately naive: the SQL fan-out is enumerated query by query, variable modifications are tried subset by
subset, the xcorr table is built densely.  Python floats are IEEE doubles without FMA, ints are exact.
"""
import itertools
import math
import re

ALPHABET = "ARNDCEQGHJKMFPOSTUVWY"                      # amino_acid.rs:4-5
MONO = {"A": 71.03711, "B": 114.53495, "R": 156.10111, "N": 114.04293, "D": 115.02694, "C": 103.00919,
        "E": 129.04259, "Q": 128.05858, "G": 57.02146, "H": 137.05891, "I": 113.08406, "L": 113.08406,
        "J": 113.08406, "K": 128.09496, "M": 131.04049, "F": 147.06841, "P": 97.05276, "O": 109.0528,
        "S": 87.03203, "T": 101.04768, "U": 150.95363, "V": 99.06841, "W": 186.07931, "X": 0.0,
        "Y": 163.06333, "Z": 128.55059}               # amino_acid.rs:7-35
H2O = 18.010565                                         # neutral_loss.rs:3
PROTON = 1.007276                                       # mass/mod.rs:4


def to_int(m):                                          # mass/mod.rs:6-8 (truncation)
    return int(m * 1000000.0)


def residue_mass(c):
    return to_int(MONO.get(c, 0.0))                     # unknown -> X -> 0 (amino_acid.rs:115)


def sequence_weight(seq):                               # amino_acid.rs:130-136
    return to_int(H2O) + sum(residue_mass(c) for c in seq)


def generalize(seq):                                    # amino_acid.rs:139-141
    return seq.replace("I", "J").replace("L", "J")


def counts21(seq):                                      # peptide_interface.rs:22-28
    return [seq.count(c) for c in ALPHABET]


_TRYPSIN = re.compile(r"(?<=[KR])(?!P)")                # trypsin.rs:29


def digest(protein, mc_max, min_len, max_len):          # digest_enzym.rs:61-86
    """-> list of (generalized sequence, missed cleavages) in emission order (duplicates kept)."""
    pieces = [p for p in _TRYPSIN.split(protein)]
    if pieces and pieces[-1] == "":
        pieces.pop()
    out = []
    for i in range(len(pieces)):
        s = ""
        for mc in range(mc_max + 1):
            j = i + mc
            if j >= len(pieces):
                break
            s += pieces[j]
            if min_len <= len(s) <= max_len:
                out.append((generalize(s), mc))
    return out


def precursor_window(mz, z, lppm, uppm):                # identification.rs:203-211
    tl = mz / 1000000.0 * float(lppm)                   # utility/mod.rs:9-11
    tu = mz / 1000000.0 * float(uppm)
    zc = float(z)
    P = to_int(mz * zc - PROTON * zc)                   # mass/mod.rs:14-16
    lo = to_int((mz - tl) * zc - PROTON * zc)
    hi = to_int((mz + tu) * zc - PROTON * zc)
    return P, lo, hi


class Mods:
    """fixed / variable maps of identification_task (identification.rs:163-196): one fixed and one variable
    modification per letter, each with its position A / N / C (modification.rs:24-33)."""

    def __init__(self, mods, nvar):
        self.fix = {m.amino_acid: m.mono_mass_int for m in mods if m.is_fix}
        self.var = {m.amino_acid: m.mono_mass_int for m in mods if not m.is_fix}
        self.fix_pos = {m.amino_acid: getattr(m, "position", "A") for m in mods if m.is_fix}
        self.var_pos = {m.amino_acid: getattr(m, "position", "A") for m in mods if not m.is_fix}
        self.nvar = nvar
        self.letters = sorted(set(self.fix) | set(self.var))        # :173-178
        self.merged = dict(self.fix)
        self.merged.update(self.var)                                # variable overrides fixed (:190-196)


class SlotPeptide:
    """ModifiedPeptide with the three modification slots the reference's add_modification_at (modified_peptide.rs:421-447)
    and set_variable_modification_at (:339-367) distinguish: `side[i]` (= modifications[i]), `nterm`
    (= n_terminus_modification, residue 0) and `cterm` (= c_terminus_modification, last residue).  Each holds None,
    ('fix', delta) or ('var', delta).  from_string = add_modification_at of the letter's fixed modification at every index."""

    def __init__(self, mods, seq):
        self.mods, self.seq = mods, seq
        self.w = to_int(H2O)
        self.side = [None] * len(seq)
        self.nterm = self.cterm = None
        for i, c in enumerate(seq):
            self.w += residue_mass(c)
            if c in mods.fix:
                self._add(i, "fix", mods.fix_pos[c], mods.fix[c])

    def _add(self, i, kind, pos, delta):
        last = len(self.seq) - 1
        if i == 0 and pos == "N":
            if self.nterm is not None:
                return False                                        # AlreadyFixModificationInPlace (:327)
            self.nterm = (kind, delta)
        elif i == last and pos == "C":
            if self.cterm is not None:
                return False                                        # :311
            self.cterm = (kind, delta)
        elif pos == "A":
            if self.side[i] is not None:
                return False                                        # :350
            self.side[i] = (kind, delta)
        else:
            return False                                            # a terminal modification away from its terminus: no effect (:345-366)
        self.w += delta
        return True

    def remove_all_variable(self):                                  # :369-401
        if self.nterm and self.nterm[0] == "var":
            self.w -= self.nterm[1]
            self.nterm = None
        if self.cterm and self.cterm[0] == "var":
            self.w -= self.cterm[1]
            self.cterm = None
        for i, s in enumerate(self.side):
            if s and s[0] == "var":
                self.w -= s[1]
                self.side[i] = None

    def set_variable(self, i):
        c = self.seq[i]
        return self._add(i, "var", self.mods.var_pos[c], self.mods.var[c])

    def var_mask(self):
        m = 0
        for i, s in enumerate(self.side):
            if s and s[0] == "var":
                m |= 1 << i
        if self.nterm and self.nterm[0] == "var":
            m |= 1
        if self.cterm and self.cterm[0] == "var":
            m |= 1 << (len(self.seq) - 1)
        return m


def fanout_queries(mods, P, lo, hi):                    # identification.rs:214-222, 374-403
    """The literal list of (lo', hi', [counts]) the reference turns into SQL queries."""
    K = {a: int(P / (residue_mass(a) + mods.merged[a])) for a in mods.letters}   # i64 division, positive operands
    K = {a: ((k + 2 ** 15) % 2 ** 16) - 2 ** 15 for a, k in K.items()}           # `as i16`
    results = []

    def rec(tol, idx, combo):
        if idx >= len(mods.letters):
            return
        a = mods.letters[idx]
        for cnt in range(0, K[a]):
            mm = cnt * mods.merged[a]
            new = (tol[0] - mm, tol[1] - mm)
            if tol[0] > 0:
                combo.append(cnt)
                if idx < len(K) - 1:
                    rec(new, idx + 1, combo)
                else:
                    results.append((new[0], new[1], list(combo)))
                combo.pop()

    rec((lo, hi), 0, [])
    return results


def n_choose_k_masks(d, k):                             # n_choose_k.rs:12-49: descending masks, MSB <-> first item
    mask = 2 ** d - 2 ** (d - k)
    lo = 2 ** k - 1
    while mask >= lo:
        if bin(mask).count("1") == k:
            yield [i for i, ch in enumerate(format(mask, "0%db" % d)) if ch == "1"]
        mask -= 1


def modified_peptide_filter(mods, seq, lo, hi):         # identification.rs:242-257; modified_peptide.rs:118-159,512-543
    """-> (accepted, weight, var position mask)"""
    mp = SlotPeptide(mods, seq)
    if lo <= mp.w <= hi:
        return True, mp.w, 0
    positions = [i for i, c in enumerate(seq) if c in mods.var]
    for n in range(1, mods.nvar + 1):
        if n > len(positions):
            continue
        for chosen in n_choose_k_masks(len(positions), n):
            mp.remove_all_variable()
            for b in chosen:
                mp.set_variable(positions[b])                       # errors -> continue 'positions
            if lo <= mp.w <= hi:
                return True, mp.w, mp.var_mask()
    return False, mp.w, 0


def candidates_sql(mods, peptides, P, lo, hi):
    """peptides: list of (sequence, weight, counts21).  Runs every fan-out query as a table scan, then the
    ModifiedPeptide filter; targets are deduplicated by sequence (fasta_entry.rs:30-44).
    -> dict peptide ordinal -> (weight, mask)"""
    col = {a: ALPHABET.index(a) for a in mods.letters}
    found = {}
    for qlo, qhi, cnts in fanout_queries(mods, P, lo, hi):
        for p, (seq, wt, c21) in enumerate(peptides):
            if not (qlo <= wt <= qhi):
                continue
            if any(c21[col[a]] != k for a, k in zip(mods.letters, cnts)):
                continue
            ok, w, mask = modified_peptide_filter(mods, seq, lo, hi)
            if ok:
                found[p] = (w, mask)
    return found


# ------------------------------------------------------------------------------------------------
# scoring (builder-defined, see oracle header): dense restatement
# ------------------------------------------------------------------------------------------------
def xcorr_table(peak_mz, peak_int, P, w, min_peaks=10):
    """-> dict bin -> T[bin] (only non-zero entries) or None if the spectrum is not scored."""
    bins, raw = [], []
    for mz, I in zip(peak_mz, peak_int):
        mz, I = float(mz), float(I)
        if not (I > 0.0) or not (mz > 0.0) or not (mz < 1.0e7):
            continue
        mzint = int(mz * 1000000.0)
        if not (0 < mzint < P + 50000000):
            continue
        bins.append(mzint // w + 1)
        raw.append(math.sqrt(I))
    if len(bins) < min_peaks or not bins:
        return None
    hbin = max(bins)
    wsize = hbin // 10 + 1
    winmax = [0.0] * 10
    for b, r in zip(bins, raw):
        winmax[b // wsize] = max(winmax[b // wsize], r)
    thr = 0.05 * max(raw)
    y = {}
    for b, r in zip(bins, raw):
        if not (r > thr):
            continue
        q = int(r * (50.0 / winmax[b // wsize]) * 65536.0 + 0.5)
        y[b] = max(y.get(b, 0), q)
    T = {}
    for b, q in y.items():
        for o in range(-75, 76):
            T[b + o] = T.get(b + o, 0) - q
        T[b] += 151 * q
    return T


def score(mods, T, seq, mask, z, w, max_frag_charge=3):
    if T is None or len(seq) < 2:
        return 0
    nch = min(max(z - 1, 1), max_frag_charge)
    m = []
    last = len(seq) - 1
    for i, c in enumerate(seq):
        v = residue_mass(c)
        if c in ALPHABET:
            fp = getattr(mods, "fix_pos", {}).get(c, "A")
            if c in mods.fix and (fp == "A" or (fp == "N" and i == 0) or (fp == "C" and i == last)):
                v += mods.fix[c]
            if (mask >> i) & 1:
                v += mods.var.get(c, 0)
        m.append(v)
    total, bsum, raw = sum(m), 0, 0
    for k in range(1, len(seq)):
        bsum += m[k - 1]
        ysum = total - bsum + to_int(H2O)
        for c in range(1, nch + 1):
            raw += T.get((bsum + c * 1007276) // (c * w) + 1, 0)
            raw += T.get((ysum + c * 1007276) // (c * w) + 1, 0)
    return raw


# ------------------------------------------------------------------------------------------------
# exhaustive decoys (builder-defined): brute force over all sequences up to a small length
# ------------------------------------------------------------------------------------------------
def exhaustive_bruteforce(mods, lo, hi, max_len):
    """All sequences over the 21 letters with 1 <= len <= max_len and fixed-mod weight in [lo, hi], ordered by
    (count vector ascending lexicographically, sequence of alphabet indices ascending)."""
    mp = [residue_mass(c) + mods.fix.get(c, 0) for c in ALPHABET]
    out = []
    for L in range(1, max_len + 1):
        for tup in itertools.product(range(21), repeat=L):
            w = to_int(H2O) + sum(mp[a] for a in tup)
            if lo <= w <= hi:
                cnt = [0] * 21
                for a in tup:
                    cnt[a] += 1
                out.append((cnt, tup, w))
    out.sort(key=lambda t: (t[0], t[1]))
    return [("".join(ALPHABET[a] for a in tup), w) for _, tup, w in out]
