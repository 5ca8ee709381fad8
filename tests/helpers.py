"""Shared graph generators for the tests (seeded, numpy only)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def powerlaw_edges(n, e, seed, alpha=1.2):
    """Directed multigraph with a heavy-tailed in/out degree (Zipf-like endpoint choice)."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, n + 1) ** alpha
    w /= w.sum()
    perm = rng.permutation(n)
    src = perm[rng.choice(n, size=e, p=w)]
    dst = perm[rng.choice(n, size=e, p=w)]
    return src.astype(np.int64), dst.astype(np.int64)


def uniform_edges(n, e, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n, e, dtype=np.int64), rng.integers(0, n, e, dtype=np.int64)
