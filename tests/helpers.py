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


# ---- a minimal tf.Example / TFRecord writer (fixtures for the sampler component) -------------------------------
def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def tf_example(features: dict) -> bytes:
    """{name: int | float | list of ints | list of floats} -> serialized tf.Example (Int64List / FloatList, packed)."""
    entries = b""
    for name, val in features.items():
        vals = list(val) if isinstance(val, (list, tuple, np.ndarray)) else [val]
        if all(isinstance(v, (int, np.integer)) for v in vals):
            feat = _ld(3, _ld(1, b"".join(_varint(int(v)) for v in vals)))
        else:
            feat = _ld(2, _ld(1, np.asarray(vals, dtype="<f4").tobytes()))
        entries += _ld(1, _ld(1, name.encode()) + _ld(2, feat))
    return _ld(1, entries)


def tfrecord_bytes(records) -> bytes:
    from gigl_b200 import sample_io as sio

    out = bytearray()
    for r in records:
        head = len(r).to_bytes(8, "little")
        out += head + sio.crc32c_masked(head).to_bytes(4, "little") + r + sio.crc32c_masked(r).to_bytes(4, "little")
    return bytes(out)
