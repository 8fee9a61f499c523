"""maxdecoy -- host side of the B200-native MaxDecoy identification hot path."""
from . import _abi
from ._abi import ALPHABET, DECOY_EXHAUSTIVE, DECOY_PERMUTE_TARGET, DECOY_REFERENCE_RANDOM, DECOY_STORED, VARMOD_EXPANDED, VARMOD_REFERENCE
from .api import Engine, MaxDecoyError, Modification, SearchParams, Spectra, load, mass, pack_proteins
from . import mzml, outputs, parallel, pgexport, synth

__all__ = ["Engine", "MaxDecoyError", "Modification", "SearchParams", "Spectra", "load", "mass", "pack_proteins",
           "mzml", "outputs", "parallel", "pgexport", "synth", "ALPHABET", "DECOY_REFERENCE_RANDOM", "DECOY_EXHAUSTIVE", "DECOY_PERMUTE_TARGET", "DECOY_STORED", "VARMOD_REFERENCE", "VARMOD_EXPANDED", "_abi"]
