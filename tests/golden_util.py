"""Helpers to load golden fixtures (tests/golden/*.npz) and replay their inputs."""
import json
import os

import numpy as np

from tests.golden.cases import CASES, case_inputs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def load_case(name):
    """-> (readset, dct, chunks, bed, golden npz) with the input checksum verified."""
    if name not in _cache:
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        rs, dct, regions, cpu, bed = case_inputs(name)
        assert rs.checksum() == str(g["input_checksum"]), (
            "synthetic inputs of golden case %r drifted; regenerate with tests/golden/make_golden.py" % name)
        chunks = json.loads(str(g["chunks_json"]))
        _cache[name] = (rs, dct, chunks, bed, g)
    return _cache[name]


def golden_chunk(g, ci):
    return {k: g["c%d_%s" % (ci, k)] for k in ("pos", "ref", "mat", "dp", "freq", "depth", "fwd", "rev")}


def available_cases():
    return [n for n in CASES if os.path.exists(os.path.join(GOLDEN_DIR, n + ".npz"))]
