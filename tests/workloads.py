"""Shared seeded workloads for the parity tests (small enough for the oracle to finish in seconds)."""
import functools
import json
import os

import numpy as np

from maxdecoy import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=None)
def p77377():
    with open(os.path.join(GOLDEN, "p77377_trypsin.json")) as fh:
        return json.load(fh)


@functools.lru_cache(maxsize=None)
def proteins(n, seed=20260101):
    return tuple(synth.synthetic_proteins(n, seed))


@functools.lru_cache(maxsize=None)
def spectra(n_prot, n_spec, mc, with_ox=False, seed=7):
    mods = (synth.CAM, synth.OXM) if with_ox else (synth.CAM,)
    return synth.synthetic_spectra(list(proteins(n_prot)), n_spec, mc, mods=mods, seed=seed)


def precursors_of(engine, sp, lppm=10, uppm=10):
    out = []
    for i in range(len(sp)):
        P, lo, hi = engine.precursor_window(float(sp.precursor_mz[i]), int(sp.charge[i]), lppm, uppm)
        out.append((P, lo, hi, int(sp.charge[i]), i))
    return out


def decoy_strings(table):
    raw = table["seq"].tobytes()
    so = table["seq_off"]
    return [raw[int(so[i]):int(so[i + 1])].decode() for i in range(len(so) - 1)]
