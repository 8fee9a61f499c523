"""Test-side loader of the CPU oracle (oracle/libmaxdecoy_oracle.so).  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs use this; the product never does."""
import os
import subprocess

import maxdecoy
from maxdecoy import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmaxdecoy_oracle.so")

_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        _lib = _abi.bind(ORACLE_SO)
    return _lib


def oracle_engine(n_threads=0):
    return maxdecoy.Engine(lib=oracle_lib(), n_threads=n_threads)
