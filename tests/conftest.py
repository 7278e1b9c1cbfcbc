import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.compile_lib()
    return oracle


@pytest.fixture(scope="session")
def ib():
    """The product package."""
    import ibvh_b200
    return ibvh_b200


def random_spheres(rng, n, fbytes=4, spread=6.0):
    """The reference tests' random scene: centre 6*rand(3), radius of the order of the triangle
    circumsphere of rand(3,3) (test/runtests.jl:849) — here simply U[0.2, 0.9)."""
    f = {4: np.float32, 8: np.float64}[fbytes]
    out = np.zeros(n, np.dtype([("x", f, 3), ("r", f)]))
    out["x"] = (spread * rng.random((n, 3))).astype(f)
    out["r"] = (0.2 + 0.7 * rng.random(n)).astype(f)
    return out


def sorted_pairs(c):
    """Structured IndexPair array or (k,2) int array -> sorted list of tuples."""
    if c.dtype.names:
        a = np.stack([c["a"].astype(np.int64), c["b"].astype(np.int64)], axis=1)
    else:
        a = np.asarray(c, np.int64).reshape(-1, 2)
    if len(a) == 0:
        return a.reshape(0, 2)
    order = np.lexsort((a[:, 1], a[:, 0]))
    return a[order]


def pairs_list(c):
    return [tuple(int(v) for v in row) for row in (np.stack([c["a"], c["b"]], axis=1) if c.dtype.names else c)]
