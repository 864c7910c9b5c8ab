"""Shared fixtures. GPU tests are marked `gpu`; everything else runs on the CPU-only container."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_SEARCH_CASES = ["col3_k4", "one_k3", "col2_k2", "two_k5"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def key64():
    with open(os.path.join(GOLDEN, "key.bin"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.lib()   # builds liboracle.so on first use
    return oracle


def read_patterns(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [l.upper() for l in lines]   # the CLI upper-cases patterns (Main.cpp:1056-1059)


@pytest.fixture(scope="session")
def golden():
    """name -> dict(index bytes, patterns, reference results)"""
    from oracle import oracle
    out = {}
    for name in GOLDEN_SEARCH_CASES + ["col2_k4_nomarks"]:
        base = os.path.join(GOLDEN, name)
        with open(base + ".efm", "rb") as f:
            d = {"index": f.read(), "patterns": read_patterns(base + ".pat"), "base": base}
        if os.path.exists(base + ".res"):
            d["res"] = oracle.read_result_file(base + ".res")
        out[name] = d
    return out


def pack(patterns):
    offs = np.zeros(len(patterns) + 1, dtype=np.uint64)
    if patterns:
        offs[1:] = np.cumsum([len(p) for p in patterns])
    return np.frombuffer(b"".join(patterns), dtype=np.uint8), offs
