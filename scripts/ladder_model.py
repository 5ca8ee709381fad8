"""Numpy model of the sampler's hash ladder (gigl_b200/csrc/khop_sample.cu: lad_plan / select_ladder / ladder_*_kernel):
same level choice, block arithmetic, thresholds and fallbacks as the CUDA code, checked against the oracle's permutation
on random windows.  A design check that runs without a GPU; the CUDA path itself is held to the oracle by
tests/test_gpu_sampler.py."""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle as O  # noqa: E402

LEVELS = 25


def ordered_keys(x):
    return (O.np_xxh64_int(x.astype(np.int32)).astype(np.uint64)) ^ np.uint64(1 << 63)


def build(limit):
    x = np.arange(limit, dtype=np.int64)
    key = ordered_keys(x)
    tk0 = (key >> np.uint64(32)).astype(np.uint32)
    lad = {0: tk0}
    for j in range(1, LEVELS + 1):
        sel = key < np.uint64(1 << (64 - j))
        xs = x[sel]
        tk = (key[sel] >> np.uint64(32 - j)).astype(np.uint64)
        ent = (tk << np.uint64(32)) | xs.astype(np.uint64)
        nb = (limit >> j) + 2
        counts = np.bincount((xs >> j) + 1, minlength=nb)
        bs = np.cumsum(counts).astype(np.uint32)  # bs[b] = entries of blocks < b
        # order inside a block is free: shuffle it to prove the point
        rng = np.random.default_rng(j)
        order = np.lexsort((rng.random(len(xs)), xs >> j))
        lad[j] = (bs, ent[order])
    return lad, key


def select(lad, key, base, s, f, stats):
    """returns window positions (1-based) of the f smallest keys in (key, pos) order, or None (-> streaming scan)"""
    limit = len(lad[0])
    if s == 0 or base + s >= limit:
        return None
    lo, hi = base + 1, base + s
    wide = 0 if f <= 16 else 1  # fanouts over 16 read a level with twice the entries
    if s < (64 << wide):
        lvl, beg, end = 0, lo, hi + 1
        ent = (lad[0][beg:end].astype(np.uint64) << np.uint64(32)) | np.arange(beg, end, dtype=np.uint64)
    else:
        lvl = min(26 - (32 - int(s).bit_length()) - wide, LEVELS)
        bs, e = lad[lvl]
        beg, end = int(bs[lo >> lvl]), int(bs[(hi >> lvl) + 1])
        ent = e[beg:end]
    stats["fetched"].append(end - beg)
    x = (ent & np.uint64(0xFFFFFFFF)).astype(np.int64)
    tk = (ent >> np.uint64(32)).astype(np.int64)
    need = min(s, f)
    inwin = (x - lo >= 0) & (x - lo < s)
    if need < 32:
        m = min(48.0, f + max(14.0, 1.2 * f))
        r = np.float32(m) * np.float32(1 << lvl) / np.float32(s)
        all_ = not (r < 1.0)
        t = 0 if all_ else int(np.float32(r) * np.float32(4294967296.0))
        tmax = 0xFFFFFFFF if all_ else (t - 1 if t else 0)
        cand = inwin & (tk <= tmax)
        c = int(cand.sum())
        if need <= c <= 64:
            nb = tmax.bit_length()
            shift = max(0, nb - 26)
            w = tk[cand] >> shift
            order = np.argsort(w, kind="stable")
            ws = w[order]
            k = min(need + 1, len(ws))
            tie = bool((ws[1:k] == ws[:k - 1]).any())
            if not tie:
                stats["fast"] += 1
                return x[cand][order][:need] - base
    cand = inwin
    c = int(cand.sum())
    if c < need or c > 160:
        return None
    stats["exact"] += 1
    xs = x[cand]
    ks = key[xs]
    order = np.lexsort((xs, ks))
    return xs[order][:need] - base


def main():
    limit = (1 << 21) + 50_000
    lad, key = build(limit)
    rng = np.random.default_rng(0)
    sizes = np.concatenate([rng.integers(1, 130, 600), (2 ** rng.uniform(6, 20.9, 1400 if len(sys.argv) > 1 else 400)).astype(np.int64),
                            np.array([63, 64, 65, 127, 128, 4095, 4096, 91701, 1 << 20])])
    for f in (15, 10, 5, 31, 32):
        stats = {"fast": 0, "exact": 0, "fetched": []}
        n_none = 0
        for s in sizes:
            s = int(s)
            base = int(rng.integers(0, limit - s - 1))
            got = select(lad, key, base, s, f, stats)
            ref = O.np_perm(s, base, 0)[:f] + 1  # positions, (key, pos) order; np_perm hashes i + internal_seed + seed
            if got is None:
                n_none += 1
                continue
            assert np.array_equal(got, ref), (f, s, base, got, ref)
        fetched = np.array(stats["fetched"])
        print(f"f = {f} ok: fast {stats['fast']}, exact {stats['exact']}, streamed {n_none}; entries fetched per row: mean "
              f"{fetched.mean():.1f}, p99 {np.percentile(fetched, 99):.0f}, max {fetched.max()}")


if __name__ == "__main__":
    main()
