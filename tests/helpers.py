"""Shared graph generators for the tests (seeded, numpy only)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def rel_err(got, ref64, ref32=None, rtol=1e-5, what=""):
    """The float parity bar of north_star ("within 1e-5 rel on fp32 embeddings"), asserted ELEMENT-WISE against the fp64
    oracle: |got - ref64| <= 1e-5 * |ref64| + atol for every element, the worst element reported on failure.

    atol: an element that cancels to ~0 cannot be relatively exact in fp32 (a K = 200-term dot product of O(1) terms carries
    ~3e-6 of absolute error), and a fixed 1e-6 fails the REFERENCE's own arithmetic - torch's fp32 index_add_ + F.linear on
    the (1000, 20000, 100, 47) case of test_gpu_aggregate has 3 of 47000 elements over 1e-5 |ref| + 1e-6 (worst 3.9e-6 at
    |ref| ~ 0.01).  So the absolute term is calibrated on that arithmetic: with `ref32` (the fp32 oracle's result for the
    same inputs) atol = 4 x its worst absolute error, never below 2e-6 - "no element further from fp64 than four times the
    reference's own worst fp32 rounding, plus 1e-5 relative" (a 3xTF32 product carries 2^-21 of relative error where an
    fp32 product carries 2^-24, so the projection's rounding is a few times fp32's; measured worst element 2.7e-6 where
    the fp32 oracle's is 1.0e-6); without it atol = 1e-5 * mean|ref64|.
    Returns the max-norm error max|got - ref| / max(1, max|ref|) that the tests also bound by 1e-5."""
    got = np.asarray(got, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    assert got.shape == ref64.shape, (got.shape, ref64.shape)
    if got.size == 0:
        return 0.0
    d = np.abs(got - ref64)
    if ref32 is not None:
        atol = max(2e-6, 4.0 * float(np.abs(np.asarray(ref32, dtype=np.float64) - ref64).max()))
    else:
        atol = rtol * float(np.abs(ref64).mean())
    slack = d - (rtol * np.abs(ref64) + atol)
    if (slack > 0).any():
        w = np.unravel_index(int(np.argmax(slack)), d.shape)
        raise AssertionError(f"{what or 'embedding'} element {w}: got {got[w]!r}, fp64 reference {ref64[w]!r}, |diff| {d[w]:.3e} > "
                             f"{rtol:g} * |ref| + {atol:.3g} ({int((slack > 0).sum())} of {d.size} elements over the bound)")
    return float(d.max() / max(1.0, np.abs(ref64).max()))


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
