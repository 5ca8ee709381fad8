"""Seeded synthetic inputs for bench.py and the tests.

Everything is counter-based integer arithmetic (a splitmix64-style mixer of the edge index), so
the numpy path (CPU, tests) and the torch path (device, bench) produce the SAME graph for the
same arguments - the GPU box and the dev container agree without shipping data.

This is input generation, not part of the product path: torch / numpy are used freely here.
"""
from __future__ import annotations

import numpy as np

GEN_SEED = 20260101  # generator seed fixed in SURVEY.md section 8(d)
RMAT_ABCD = (0.57, 0.19, 0.19, 0.05)

_M64 = (1 << 64) - 1
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB
_GOLD = 0x9E3779B97F4A7C15


def _s64(v: int) -> int:
    """python int -> the same 64 bits as a signed int64 value."""
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _mix64(z, xp):
    """splitmix64 finaliser on int64 tensors/arrays with wrapping multiply and logical shifts."""
    z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * _s64(_C1)
    z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * _s64(_C2)
    z = z ^ ((z >> 31) & ((1 << 33) - 1))
    return z


def _rmat_chunk(xp, idx, scale, seed, n_nodes, abcd):
    """idx: int64 edge indices -> (src, dst) int64 in [0, n_nodes)."""
    a, b, c, _ = abcd
    ta = int(a * (1 << 24))
    tb = int((a + b) * (1 << 24))
    tc = int((a + b + c) * (1 << 24))
    src = idx * 0
    dst = idx * 0
    base = idx * _s64(_GOLD) + _s64(seed * 0x2545F4914F6CDD1D)
    for level in range(scale):
        r = (_mix64(base + _s64((level + 1) * 0xD6E8FEB86659FD93), xp) >> 40) & ((1 << 24) - 1)
        sbit = (r >= tb).to(idx.dtype) if xp is None else (r >= tb).astype(np.int64)
        in_b = (r >= ta) & (r < tb)
        in_d = r >= tc
        dbit = (in_b | in_d).to(idx.dtype) if xp is None else (in_b | in_d).astype(np.int64)
        src = src * 2 + sbit
        dst = dst * 2 + dbit
    # scramble ids inside the 2^scale space (odd multiplier = bijection), then fold into [0, n_nodes)
    mask = (1 << scale) - 1
    src = ((src * 0x9E3779B1 + 0x7F4A7C15) & mask) % n_nodes
    dst = ((dst * 0x85EBCA6B + 0x165667B1) & mask) % n_nodes
    return src, dst


def rmat_edges_numpy(n_nodes: int, n_edges: int, seed: int = GEN_SEED, abcd=RMAT_ABCD, start: int = 0):
    """RMAT/Kronecker power-law edge list (numpy, int64)."""
    scale = max(1, int(np.ceil(np.log2(max(n_nodes, 2)))))
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + n_edges, dtype=np.int64)
        return _rmat_chunk(np, idx, scale, seed, n_nodes, abcd)


def rmat_edges_torch(n_nodes: int, n_edges: int, device, seed: int = GEN_SEED, abcd=RMAT_ABCD, chunk: int = 1 << 25):
    """Same edge list generated on `device`; returns int32 (src, dst) tensors."""
    import torch

    scale = max(1, int(np.ceil(np.log2(max(n_nodes, 2)))))
    src = torch.empty(n_edges, dtype=torch.int32, device=device)
    dst = torch.empty(n_edges, dtype=torch.int32, device=device)
    for s in range(0, n_edges, chunk):
        m = min(chunk, n_edges - s)
        idx = torch.arange(s, s + m, dtype=torch.int64, device=device)
        a, b = _rmat_chunk(None, idx, scale, seed, n_nodes, abcd)
        src[s:s + m] = a.to(torch.int32)
        dst[s:s + m] = b.to(torch.int32)
    return src, dst


def features_torch(n: int, F: int, device, seed: int = GEN_SEED):
    """[n, F] fp32 ~ N(0,1) on `device` (torch generator; values differ from features_numpy)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randn(n, F, device=device, dtype=torch.float32, generator=g)


def sage_weights(rng: np.random.Generator, dims):
    """[(Wl [O,F], bl [O], Wr [O,F]), ...] for consecutive dims, Glorot-ish scale."""
    out = []
    for F, O in zip(dims[:-1], dims[1:]):
        s = 1.0 / np.sqrt(F)
        out.append(((rng.standard_normal((O, F)) * s).astype(np.float32),
                    (rng.standard_normal(O) * 0.1).astype(np.float32),
                    (rng.standard_normal((O, F)) * s).astype(np.float32)))
    return out
