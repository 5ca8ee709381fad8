// khop_sample.cu - k-hop rooted-neighbourhood index sampling for sm_100a.
//
// Replaces sampleOnehopSrcNodesUniformly / sampleTwohopSrcNodesUniformly
// (scala/subgraph_sampler/src/main/scala/libs/task/pureSpark/SGSPureSparkV1Task.scala:313-494) and
// SamplingStrategy.hashBasedUniformPermutation (libs/task/SamplingStrategy.scala:16-82):
// the GROUP BY / array_sort / xxhash64 / slice pipeline becomes one warp per frontier slot that
// hashes the row's index window and keeps the `fanout` smallest keys in registers.
//
// Key observation: the permutation key of position i depends only on (i, internal_seed, seed),
// never on the neighbour id stored at i.  So selection is pure integer ALU work on the window
// x = base+1 .. base+size (base = internal_seed + seed, int32 wrapping) and memory is touched only
// for the rowptr pair and the <= fanout winning column entries.
//
// Layout in HBM: rowptr int64[n+1], col int32[E] (rows sorted ascending), padded-tree outputs
// nbr[h] int32[n_roots * prod f], cnt[h] int32[n_roots * prod f_{<h}]  (see include/gigl_b200.h).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>

#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "xxh64.cuh"

namespace gigl {

constexpr int kWarpsPerBlock = 8;
#ifndef KHOP_MIN_BLOCKS
#define KHOP_MIN_BLOCKS 8  // the kernels are latency-bound: 64 resident warps per SM (32 registers) beat 40 (46 registers) by 16 %
#endif
constexpr int kHeavyThreshold = 4096;   // rows longer than this go to the CTA-per-row pass
constexpr int kHeavyWarps = 16;

struct HopArgs {
    const int64_t* rowptr;
    const int32_t* col;
    int64_t n_nodes;
    const int32_t* roots;
    const int32_t* levels[GIGL_MAX_HOPS];  // levels[k] = nbr of hop k+1 (device), k < h-1
    int32_t fanouts[GIGL_MAX_HOPS];
    int32_t h;         // current hop, 1-based
    int32_t cur_seed;  // base_seed * (first_call_no + h - 1), int32 wrapping
    int32_t* out_nbr;  // [n_parent * f]
    int32_t* out_cnt;  // [n_parent]
    int64_t n_parent;  // parent slots at this hop
    int32_t* err;      // device error word
    int32_t* heavy_list;  // worklist of parent slots deferred to the heavy pass
    int32_t* heavy_count;
    int32_t heavy_cap;
    unsigned long long* tile_counter;  // next tile of khop_tile_kernel (zeroed per launch)
    // hash ladder (see the "hash ladder" section below); lad_limit == 0 = disabled
    const uint32_t* lad_tk0;
    const uint32_t* lad_bs[GIGL_LAD_MAX_LEVELS + 1];
    const uint64_t* lad_ent[GIGL_LAD_MAX_LEVELS + 1];
    uint64_t lad_limit;
    int32_t lad_levels;
    // stage-slot claims of the vertices this hop writes (gigl_stage_args); st_slot == nullptr = none
    int32_t* st_slot;
    int32_t* st_list;
    int32_t* st_ctr;
};

// First writer of vertex v into the tree claims its stage slot (the slot NUMBER is assigned by the caller, in bulk).
__device__ __forceinline__ bool stage_try_claim(const HopArgs& a, int32_t v) {
    return v >= 0 && a.st_slot[v] == kStageAbsent && atomicCAS(a.st_slot + v, kStageAbsent, kStagePending) == kStageAbsent;
}

// Sorted (ascending) best-`f` list held by one warp: position p lives in lane p%32, register p/32.
template <int KPL>
struct WarpTopK {
    uint64_t key[KPL];
    int32_t idx[KPL];
    uint64_t kth;  // key at position f-1 (kKeyInf until f entries were inserted)
    bool empty;    // nothing inserted yet (warp-uniform)

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            key[k] = kKeyInf;
            idx[k] = 0;
        }
        kth = kKeyInf;
        empty = true;
    }

    // Insert (ck, ci), warp-uniform arguments, ck < kth.
    __device__ __forceinline__ void insert(uint64_t ck, int32_t ci, int f, int lane) {
        empty = false;
        int pos = 0;
#pragma unroll
        for (int k = 0; k < KPL; ++k) pos += __popc(__ballot_sync(0xffffffffu, key[k] < ck));
#pragma unroll
        for (int k = KPL - 1; k >= 0; --k) {
            uint64_t upk = __shfl_up_sync(0xffffffffu, key[k], 1);
            int32_t upi = __shfl_up_sync(0xffffffffu, idx[k], 1);
            if (k > 0) {
                uint64_t pk = __shfl_sync(0xffffffffu, key[k - 1], 31);
                int32_t pi = __shfl_sync(0xffffffffu, idx[k - 1], 31);
                if (lane == 0) {
                    upk = pk;
                    upi = pi;
                }
            }
            const int p = k * 32 + lane;
            if (p > pos) {
                key[k] = upk;
                idx[k] = upi;
            } else if (p == pos) {
                key[k] = ck;
                idx[k] = ci;
            }
        }
        refresh_kth(f);
    }

    // insert() places a key BEFORE the equal keys already held; this one places it after them, so that candidates offered in
    // ascending position order end up ordered by (key, position) - the weighted ops' tie rule.
    __device__ __forceinline__ void insert_stable(uint64_t ck, int32_t ci, int f, int lane) {
        empty = false;
        int pos = 0;
#pragma unroll
        for (int k = 0; k < KPL; ++k) pos += __popc(__ballot_sync(0xffffffffu, key[k] <= ck));
#pragma unroll
        for (int k = KPL - 1; k >= 0; --k) {
            uint64_t upk = __shfl_up_sync(0xffffffffu, key[k], 1);
            int32_t upi = __shfl_up_sync(0xffffffffu, idx[k], 1);
            if (k > 0) {
                uint64_t pk = __shfl_sync(0xffffffffu, key[k - 1], 31);
                int32_t pi = __shfl_sync(0xffffffffu, idx[k - 1], 31);
                if (lane == 0) {
                    upk = pk;
                    upi = pi;
                }
            }
            const int p = k * 32 + lane;
            if (p > pos) {
                key[k] = upk;
                idx[k] = upi;
            } else if (p == pos) {
                key[k] = ck;
                idx[k] = ci;
            }
        }
        refresh_kth(f);
    }

    __device__ __forceinline__ void refresh_kth(int f) {
        const int p = f - 1;
        uint64_t v = kKeyInf;
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            uint64_t t = __shfl_sync(0xffffffffu, key[k], p & 31);
            if ((p >> 5) == k) v = t;
        }
        kth = v;
    }

    // First 32 candidates of an EMPTY list (KPL == 1 fast path): a 32-lane bitonic sort of
    // (key, idx) instead of 32 serial insertions.  k = kKeyInf marks an absent candidate.
    __device__ __forceinline__ void seed_sorted32(uint64_t k, int32_t i, int f, int lane) {
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const uint64_t ok = __shfl_xor_sync(0xffffffffu, k, stride);
                const int32_t oi = __shfl_xor_sync(0xffffffffu, i, stride);
                const bool up = ((lane & size) == 0);           // ascending block?
                const bool lower = ((lane & stride) == 0);      // I hold the lower position of the pair
                const bool take_min = (up == lower);
                const bool other_less = ok < k;
                if (take_min == other_less) {
                    k = ok;
                    i = oi;
                }
            }
        }
        key[0] = k;
        idx[0] = i;
        empty = false;
        refresh_kth(f);
    }

    // Scan window positions [i_begin, i_end] (1-based, inclusive) with stride `stride` chunks of 32.
    __device__ __forceinline__ void scan(int64_t i_first_chunk, int64_t size, int64_t chunk_stride, uint32_t base,
                                         int f, int lane) {
        for (int64_t c0 = i_first_chunk; c0 < size; c0 += chunk_stride) {
            const int64_t i = c0 + lane + 1;  // 1-based index as F.sequence(1, size)
            uint64_t k = kKeyInf;
            if (i <= size) k = ordered_key((int32_t)(base + (uint32_t)i));
            if (KPL == 1 && empty) {  // warp-uniform
                seed_sorted32(k, (int32_t)i, f, lane);
                continue;
            }
            uint32_t cand = __ballot_sync(0xffffffffu, k < kth);
            while (cand) {
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                const uint64_t ck = __shfl_sync(0xffffffffu, k, src);
                if (ck < kth) insert(ck, (int32_t)(c0 + src + 1), f, lane);
            }
        }
    }
};

// Resolves parent vertex, path-id sum and sibling multiplicity for one parent slot (warp-uniform).
// Returns false if the slot produces no group (empty parent or a non-first duplicate).
__device__ __forceinline__ bool resolve_parent(const HopArgs& a, int64_t pslot, int lane, int32_t& v, uint32_t& ssum,
                                               int32_t& mult) {
    mult = 1;
    if (a.h == 1) {
        v = a.roots[pslot];
        ssum = (uint32_t)v;
        return true;
    }
    const int32_t* prev = a.levels[a.h - 2];
    v = prev[pslot];
    if (v < 0) return false;
    ssum = (uint32_t)v;
    int64_t s = pslot;
    for (int hh = a.h - 1; hh >= 1; --hh) {
        s /= a.fanouts[hh - 1];
        ssum += (uint32_t)(hh == 1 ? a.roots[s] : a.levels[hh - 2][s]);
    }
    const int fp = a.fanouts[a.h - 2];
    const int64_t sib0 = (pslot / fp) * fp;
    const int me = (int)(pslot - sib0);
    int m = 0;
    bool first = true;
    for (int j0 = 0; j0 < fp; j0 += 32) {
        const int j = j0 + lane;
        const int32_t sv = (j < fp) ? prev[sib0 + j] : -1;
        const uint32_t eq = __ballot_sync(0xffffffffu, sv == v);
        m += __popc(eq);
        if (me >= j0) {
            const int upto = me - j0;  // bits strictly below my own position in this word
            const uint32_t below = upto >= 32 ? eq : (eq & ((1u << upto) - 1u));
            if (below) first = false;
        }
    }
    mult = m;
    return first;
}

template <int KPL>
__device__ __forceinline__ void write_result(const HopArgs& a, int64_t pslot, const WarpTopK<KPL>& best, int64_t size,
                                             int64_t row_begin, int32_t mult, int lane) {
    const int f = a.fanouts[a.h - 1];
    const int n_out = (int)(size < f ? size : f);
    int32_t* o = a.out_nbr + pslot * f;
#pragma unroll
    for (int k = 0; k < KPL; ++k) {
        const int p = k * 32 + lane;
        int32_t r = -1;
        if (p < f) {
            if (p < n_out) {
                const int64_t j = (int64_t)(best.idx[k] - 1) / mult;  // sorted(m copies)[q] = row[q / m]
                r = __ldg(a.col + row_begin + j);
            }
            o[p] = r;
        }
        if (a.st_slot != nullptr) {  // warp-uniform; rows of this path are rare (long rows, fanouts over 32)
            const bool won = p < f && stage_try_claim(a, r);
            const uint32_t m = __ballot_sync(0xffffffffu, won);
            if (m) {
                int s0 = 0;
                if (lane == __ffs(m) - 1) s0 = atomicAdd(a.st_ctr, __popc(m));
                s0 = __shfl_sync(0xffffffffu, s0, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
                if (won) {
                    a.st_list[s0] = r;
                    a.st_slot[r] = s0;
                }
            }
        }
    }
    if (lane == 0) a.out_cnt[pslot] = n_out;
}

// Exact top-f of one row's window by streaming every key through the sorted list (the path every other one falls
// back to).
template <int KPL>
__device__ __forceinline__ void select_row(const HopArgs& a, WarpTopK<KPL>& best, int64_t size, uint32_t base, int f,
                                           int lane) {
    best.scan(0, size, 32, base, f, lane);
}

// ---- threshold selection (fanout <= 32) ------------------------------------------------------
// The keys are uniform 64-bit values, so the f-th smallest of `size` keys sits near (f / size) * 2^64.
// One pass collects every key below T = (m / size) * 2^64, m a little above f, into a per-warp
// shared-memory buffer (no serial insertions), one bitonic sort orders the <= 64 candidates, the
// first f are the answer.  If fewer than f keys fell below T, or more than the buffer holds (both
// rare by the Poisson tail), an exact path redoes the row - the result is always the exact top-f of
// the exact keys.
constexpr int kCandCap = 64;        // candidates of a threshold pass
constexpr int kCandCapWide = 160;   // candidates of the ladder's exact second attempt

struct PairKI {
    uint64_t k;
    int32_t i;
};

__device__ __forceinline__ void cmpx(PairKI& v, int lane, int stride, bool take_min) {
    const uint64_t ok = __shfl_xor_sync(0xffffffffu, v.k, stride);
    const int32_t oi = __shfl_xor_sync(0xffffffffu, v.i, stride);
    if (take_min == (ok < v.k)) {
        v.k = ok;
        v.i = oi;
    }
}

__device__ __forceinline__ void sort32(PairKI& v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1)
            cmpx(v, lane, stride, ((lane & size) == 0) == ((lane & stride) == 0));
}

__device__ __forceinline__ void cmpx_u32(uint32_t& v, int stride, bool take_min) {
    const uint32_t o = __shfl_xor_sync(0xffffffffu, v, stride);
    v = take_min ? min(v, o) : max(v, o);
}

__device__ __forceinline__ void sort32_u32(uint32_t& v, int lane) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) cmpx_u32(v, stride, ((lane & size) == 0) == ((lane & stride) == 0));
}

// Orders c <= 64 buffered candidates on ONE 32-bit word each: 26 significant key bits (word_of(slot), < 2^26) over the
// candidate's 6-bit buffer slot.  A 32-lane bitonic network on such words is one shuffle and one min/max per stage
// instead of three shuffles and a 64-bit compare.  On return lane l < need holds the buffer slot of the l-th smallest
// key in v0 & 63.  Two candidates agreeing on the 26 bits inside the first need + 1 positions (about 1e-5 of the rows)
// would make the order ambiguous: returns false and the row is redone by an exact path.
template <class WordOf>
__device__ __forceinline__ bool order_cands_u32(int c, int need, int lane, WordOf word_of, uint32_t& v0) {
    uint32_t v1 = 0xFFFFFFFFu;
    v0 = 0xFFFFFFFFu;
    if (lane < c) v0 = (word_of(lane) << 6) | (uint32_t)lane;
    sort32_u32(v0, lane);
    if (c > 32) {  // warp-uniform
        if (lane + 32 < c) v1 = (word_of(lane + 32) << 6) | (uint32_t)(lane + 32);
        sort32_u32(v1, lane);
        // 32 smallest of two ascending runs: min(v0[l], v1[31 - l]) is bitonic; one merge sorts it
        v0 = min(v0, __shfl_sync(0xffffffffu, v1, 31 - lane));
#pragma unroll
        for (int stride = 16; stride > 0; stride >>= 1) cmpx_u32(v0, stride, (lane & stride) == 0);
    }
    const uint32_t up = __shfl_up_sync(0xffffffffu, v0, 1);
    const bool tie = lane >= 1 && lane <= need && v0 != 0xFFFFFFFFu && (v0 >> 6) == (up >> 6);
    return !__any_sync(0xffffffffu, tie);
}

__device__ __forceinline__ bool push_cands(uint64_t* __restrict__ ck, int32_t* __restrict__ ci, int& c, bool is_cand,
                                           uint64_t k, int32_t i, int lane) {
    const uint32_t m = __ballot_sync(0xffffffffu, is_cand);
    if (m == 0) return true;
    const int n = __popc(m);
    if (c + n > kCandCap) return false;
    if (is_cand) {
        const int p = c + __popc(m & ((1u << lane) - 1u));
        ck[p] = k;
        ci[p] = i;
    }
    c += n;
    return true;
}

// Rows the ladder does not cover (hash inputs at or beyond its limit, or no ladder): keys hashed on the fly.
// Returns true with best.key[0] / best.idx[0] = the row's keys in ascending order (>= f of them, or all).
__device__ __forceinline__ bool select_threshold(const HopArgs& a, WarpTopK<1>& best, int64_t size, uint32_t base, int f,
                                                 int lane, uint64_t* ck, int32_t* ci) {
    float m = (float)f + fmaxf(14.f, 1.2f * (float)f);
    if (m > 48.f) m = 48.f;
    uint64_t T = kKeyInf;
    if ((float)size > m) T = __float2ull_rz(m / (float)size * 18446744073709551616.0f);
    int c = 0;
    __syncwarp();  // a previous row that gave up early left its candidate writes unordered against these
    for (int64_t c0 = 0; c0 < size; c0 += 32) {
        const int64_t i = c0 + lane + 1;
        uint64_t k = kKeyInf;
        if (i <= size) k = ordered_key((int32_t)(base + (uint32_t)i));
        if (!push_cands(ck, ci, c, k < T, k, (int32_t)i, lane)) return false;
    }
    const int64_t need = size < f ? size : f;
    if (c < need) return false;
    __syncwarp();
    if (need < 32) {
        int shift = 38;
        if (T != kKeyInf) {
            const int nbits = 64 - __clzll((long long)T);
            shift = nbits > 26 ? nbits - 26 : 0;
        }
        uint32_t v0;
        const bool ok = order_cands_u32(c, (int)need, lane, [&](int slot) { return (uint32_t)(ck[slot] >> shift); }, v0);
        if (ok && lane < need) {
            best.key[0] = ck[v0 & 63u];
            best.idx[0] = ci[v0 & 63u];
        }
        __syncwarp();
        return ok;
    }
    PairKI v0{kKeyInf, 0}, v1{kKeyInf, 0};
    if (lane < c) {
        v0.k = ck[lane];
        v0.i = ci[lane];
    }
    sort32(v0, lane);
    if (c > 32) {  // warp-uniform
        if (lane + 32 < c) {
            v1.k = ck[lane + 32];
            v1.i = ci[lane + 32];
        }
        sort32(v1, lane);
        const uint64_t ok = __shfl_sync(0xffffffffu, v1.k, 31 - lane);
        const int32_t oi = __shfl_sync(0xffffffffu, v1.i, 31 - lane);
        if (ok < v0.k) {
            v0.k = ok;
            v0.i = oi;
        }
#pragma unroll
        for (int stride = 16; stride > 0; stride >>= 1) cmpx(v0, lane, stride, (lane & stride) == 0);
    }
    __syncwarp();
    best.key[0] = v0.k;
    best.idx[0] = v0.i;
    return true;
}

// ---- hash ladder ------------------------------------------------------------------------------
// The permutation key of window position i is H(base + i): a row of `size` entries asks for the f smallest values of
// the FIXED sequence H(x) over the window x in [base + 1, base + size].  The sequence is indexed once per context
// (gigl_ladder, common.cuh): level j holds exactly the inputs whose key is below 2^(64 - j) - a 2^-j sample of the
// inputs - grouped by block x >> j, each as (key >> (32 - j)) << 32 | x.  A row of size s picks the level where
// s >> j is in [32, 64): the f smallest keys of its window are, but for a Poisson tail, among the level's entries
// inside the window, and those entries are ONE contiguous run of 32 .. 66 eight-byte words between two block starts.
// So a row costs two 4-byte block-start reads (lane-parallel for the 32 rows of a tile) and two or three coalesced
// 256-byte reads, whatever its length - a 91 701-entry hub row reads what a 64-entry row reads.  Rows under 64 entries
// read their 4-byte truncated keys at level 0.  Candidates are cut by the same expected-count threshold as above (on
// the truncated key: any threshold is valid, the candidates are then exactly the keys below it), ordered on 26 bits,
// and every failure (too few / too many candidates, a 26-bit tie) is redone from the level's whole run with exact
// 64-bit keys; a level with fewer than f keys inside the window (probability < 1e-4 at s >> j = 32, vanishing above)
// falls through to the streaming scan.  Bit-exact by construction: the same keys, the same (key, idx) order.
// tmax: a truncated key is a candidate iff it is <= tmax (2^32 - 1 = every entry inside the window)
__device__ __forceinline__ void lad_plan(const HopArgs& a, uint32_t base, uint32_t s, int f, int& lvl, uint32_t& beg, uint32_t& end,
                                         uint32_t& tmax) {
    lvl = -1;
    beg = end = 0;
    tmax = 0xFFFFFFFFu;
    if (a.lad_limit == 0 || s == 0 || (uint64_t)base + (uint64_t)s >= a.lad_limit) return;
    const uint32_t lo = base + 1u, hi = base + s;
    const int wide = f <= 16 ? 0 : 1;  // fanouts over 16 read a level with twice the entries (s >> j in [64, 128))
    int j = 0;
    if (s >= (64u << wide)) {
        j = 26 - __clz((int)s) - wide;  // s >> j in [32, 64)
        if (j > a.lad_levels) j = a.lad_levels;
    }
    {   // expected-count threshold on the level's truncated keys: m / (s >> j) of them are wanted
        float m = (float)f + fmaxf(14.f, 1.2f * (float)f);
        if (m > 48.f) m = 48.f;
        const float r = m * (float)(1u << j) / (float)s;
        if (r < 1.f) {
            const uint32_t t = __float2uint_rz(r * 4294967296.f);
            tmax = t ? t - 1u : 0u;
        }
    }
    if (j == 0) {
        lvl = 0;
        beg = lo;
        end = hi + 1u;
        return;
    }
    lvl = j;
    const uint32_t* bs = a.lad_bs[j];
    beg = __ldg(bs + (lo >> j));
    end = __ldg(bs + (hi >> j) + 1);
}

// entry e of the level as (truncated key) << 32 | x; ~0 past the run (never inside a window: x = 2^32 - 1)
__device__ __forceinline__ uint64_t lad_load(const HopArgs& a, int lvl, uint32_t e, uint32_t end) {
    if (e >= end) return ~0ULL;
    if (lvl == 0) return ((uint64_t)__ldg(a.lad_tk0 + e) << 32) | e;
    return __ldg(a.lad_ent[lvl] + e);
}

__device__ __forceinline__ bool push_cands32(uint32_t* __restrict__ ctk, uint32_t* __restrict__ cx, int& c, bool is_cand,
                                             uint32_t tk, uint32_t x, int lane, int cap) {
    const uint32_t m = __ballot_sync(0xffffffffu, is_cand);
    if (m == 0) return true;
    const int n = __popc(m);
    if (c + n > cap) return false;
    if (is_cand) {
        const int p = c + __popc(m & ((1u << lane) - 1u));
        ctk[p] = tk;
        cx[p] = x;
    }
    c += n;
    return true;
}

// `first` = lad_load of the run's first 32 entries (the caller has it in flight while the previous row is selected).
// Returns true with best.idx[0] of lane l < min(s, f) = window position of the l-th smallest key.
__device__ __forceinline__ bool select_ladder(const HopArgs& a, WarpTopK<1>& best, uint32_t s, uint32_t base, int f, int lane,
                                              int lvl, uint32_t beg, uint32_t end, uint32_t tmax, uint64_t first, uint32_t* ctk,
                                              uint32_t* cx) {
    const uint32_t lo = base + 1u;
    const int need = (int)(s < (uint32_t)f ? s : (uint32_t)f);
    __syncwarp();  // orders this row's candidate writes after whatever the previous row left (it may have given up early)
    if (need < 32) {
        int c = 0;
        bool ok = true;
        for (uint32_t e0 = beg; e0 < end; e0 += 32) {
            const uint64_t ent = (e0 == beg) ? first : lad_load(a, lvl, e0 + lane, end);
            const uint32_t x = (uint32_t)ent, tk = (uint32_t)(ent >> 32);
            const bool cand = (x - lo) < s && tk <= tmax;
            if (!push_cands32(ctk, cx, c, cand, tk, x, lane, kCandCap)) {
                ok = false;
                break;
            }
        }
        if (ok && c >= need) {
            __syncwarp();
            const int nb = 32 - __clz((int)tmax);  // candidates are <= tmax: their bits above its highest are zero
            const int shift = nb > 26 ? nb - 26 : 0;
            uint32_t v0;
            ok = order_cands_u32(c, need, lane, [&](int slot) { return ctk[slot] >> shift; }, v0);
            if (ok && lane < need) best.idx[0] = (int32_t)(cx[v0 & 63u] - base);
            __syncwarp();
            if (ok) return true;
        }
    }
    // second attempt: every entry of the level inside the window, exact 64-bit keys through the sorted list
    int c = 0;
    __syncwarp();
    for (uint32_t e0 = beg; e0 < end; e0 += 32) {
        const uint64_t ent = (e0 == beg) ? first : lad_load(a, lvl, e0 + lane, end);
        const uint32_t x = (uint32_t)ent;
        if (!push_cands32(ctk, cx, c, (x - lo) < s, (uint32_t)(ent >> 32), x, lane, kCandCapWide)) return false;
    }
    if (c < need) return false;
    __syncwarp();
    best.init();
    for (int p0 = 0; p0 < c; p0 += 32) {
        const int p = p0 + lane;
        uint64_t k = kKeyInf;
        int32_t i = 0;
        if (p < c) {
            const uint32_t x = cx[p];
            k = ordered_key((int32_t)x);
            i = (int32_t)(x - base);
        }
        if (best.empty) {  // warp-uniform
            best.seed_sorted32(k, i, f, lane);
            continue;
        }
        uint32_t cand = __ballot_sync(0xffffffffu, k < best.kth);
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            const uint64_t ck = __shfl_sync(0xffffffffu, k, src);
            const int32_t ci = __shfl_sync(0xffffffffu, i, src);
            if (ck < best.kth) best.insert(ck, ci, f, lane);
        }
    }
    __syncwarp();
    return true;
}

template <int KPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, KHOP_MIN_BLOCKS) khop_hop_kernel(const HopArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t pslot = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (pslot >= a.n_parent) return;
    const int f = a.fanouts[a.h - 1];
    int32_t v, mult;
    uint32_t ssum;
    const bool live = resolve_parent(a, pslot, lane, v, ssum, mult);
    int64_t size = 0, row_begin = 0;
    bool ok = live;
    if (live) {
        if (v < 0 || v >= a.n_nodes) {
            if (lane == 0) atomicExch(a.err, GIGL_E_RANGE);
            ok = false;
        } else {
            row_begin = __ldg(a.rowptr + v);
            size = (__ldg(a.rowptr + v + 1) - row_begin) * mult;
            if (size > 2147483647LL) {
                if (lane == 0) atomicExch(a.err, GIGL_E_OVERFLOW);
                ok = false;
            }
        }
    }
    if (!ok) {
        int32_t* o = a.out_nbr + pslot * f;
        for (int p = lane; p < f; p += 32) o[p] = -1;
        if (lane == 0) a.out_cnt[pslot] = 0;
        return;
    }
    const uint32_t base = ssum + (uint32_t)a.cur_seed;
    int lvl = -1;
    uint32_t e_beg = 0, e_end = 0, tmax = 0;
    if (KPL == 1) lad_plan(a, base, (uint32_t)size, f, lvl, e_beg, e_end, tmax);
    if (size > kHeavyThreshold && a.heavy_list != nullptr && lvl < 0) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(a.heavy_count, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot < a.heavy_cap) {
            if (lane == 0) a.heavy_list[slot] = (int32_t)pslot;
            return;
        }
        // worklist full: fall through and do the long row with this warp alone (still exact)
    }
    WarpTopK<KPL> best;
    best.init();
    if constexpr (KPL == 1) {
        __shared__ uint64_t s_cand[kWarpsPerBlock][kCandCapWide];
        const int w = threadIdx.x >> 5;
        uint32_t* ctk = reinterpret_cast<uint32_t*>(s_cand[w]);
        bool done = false;
        if (lvl >= 0)
            done = select_ladder(a, best, (uint32_t)size, base, f, lane, lvl, e_beg, e_end, tmax, lad_load(a, lvl, e_beg + lane, e_end), ctk,
                                 ctk + kCandCapWide);
        if (!done && (size <= 32 || !select_threshold(a, best, size, base, f, lane, s_cand[w], reinterpret_cast<int32_t*>(s_cand[w] + kCandCap)))) {
            best.init();
            select_row<KPL>(a, best, size, base, f, lane);
        }
    } else {
        select_row<KPL>(a, best, size, base, f, lane);
    }
    write_result<KPL>(a, pslot, best, size, row_begin, mult, lane);
}

// ---- weighted SamplingOps --------------------------------------------------------------------------
// TopK and RandomWeighted of subgraph_sampling_strategy.proto as the reference's Nebula translator words them
// (NebulaQueryResponseTranslator.scala:73-105): "ORDER BY <edge feature> DESC | LIMIT k", and for RandomWeighted the same
// with the feature multiplied by rand().  One warp per frontier slot streams the row's weights once; score = w (top-k) or
// w * u, u in (0, 1) drawn from the op's seeded hash permutation key of the window position (the reference's rand() is
// unseeded); the f largest scores win, ties to the lower CSR position (= lower neighbour id), NaN weights last.
__device__ __forceinline__ uint64_t desc_key_f64(double s) {  // ascending order of the result = descending order of s
    if (s != s) s = -CUDART_INF;
    s += 0.0;  // -0.0 -> +0.0
    uint64_t b = (uint64_t)__double_as_longlong(s);
    b ^= (b >> 63) ? ~0ULL : 0x8000000000000000ULL;
    return ~b;  // never kKeyInf: the largest result is that of -inf, 0xFFF0000000000000
}

template <int KPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) khop_weighted_kernel(const HopArgs a, const float* __restrict__ w, int method) {
    const int lane = threadIdx.x & 31;
    const int64_t pslot = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (pslot >= a.n_parent) return;
    const int f = a.fanouts[a.h - 1];
    int32_t v, mult;
    uint32_t ssum;
    bool ok = resolve_parent(a, pslot, lane, v, ssum, mult);  // a repeated sibling is expanded once (its first slot)
    int64_t size = 0, row_begin = 0;
    if (ok) {
        if (v < 0 || v >= a.n_nodes) {
            if (lane == 0) atomicExch(a.err, GIGL_E_RANGE);
            ok = false;
        } else {
            row_begin = __ldg(a.rowptr + v);
            size = __ldg(a.rowptr + v + 1) - row_begin;
            if (size > 2147483647LL) {
                if (lane == 0) atomicExch(a.err, GIGL_E_OVERFLOW);
                ok = false;
            }
        }
    }
    if (!ok) {
        int32_t* o = a.out_nbr + pslot * f;
        for (int p = lane; p < f; p += 32) o[p] = -1;
        if (lane == 0) a.out_cnt[pslot] = 0;
        return;
    }
    const uint32_t base = ssum + (uint32_t)a.cur_seed;
    WarpTopK<KPL> best;
    best.init();
    for (int64_t c0 = 0; c0 < size; c0 += 32) {
        const int64_t i = c0 + lane + 1;
        uint64_t k = kKeyInf;
        if (i <= size) {
            double sc = (double)__ldg(w + row_begin + i - 1);
            if (method == GIGL_SAMPLE_RANDOM_WEIGHTED) {
                const uint64_t hk = ordered_key((int32_t)(base + (uint32_t)i));
                sc *= ((double)(hk >> 12) + 0.5) * (1.0 / 4503599627370496.0);  // u = (top 52 bits + 1/2) * 2^-52, exact in fp64
            }
            k = desc_key_f64(sc);
        }
        uint32_t cand = __ballot_sync(0xffffffffu, k < best.kth);
        while (cand) {
            const int src = __ffs(cand) - 1;
            cand &= cand - 1;
            const uint64_t ck = __shfl_sync(0xffffffffu, k, src);
            if (ck < best.kth) best.insert_stable(ck, (int32_t)(c0 + src + 1), f, lane);
        }
    }
    write_result<KPL>(a, pslot, best, size, row_begin, 1, lane);
}

// ---- tile form of the hop kernel (fanout <= 32) ---------------------------------------------------
// A warp-per-row kernel is latency-bound (ncu: 41 % of the stall samples wait on global loads, DRAM 7 % busy): every
// warp walks ONE row through its dependent round trips (parent slot -> rowptr pair -> block starts -> ladder run ->
// winning col entries).  Here a warp owns a tile of 32 consecutive parent slots:
//   A. lane l resolves slot l's metadata (parent vertex, path-id sum, sibling multiplicity, rowptr pair, ladder level
//      and block starts): the same loads, but 32 rows share each round trip;
//   B. the rows are selected one after the other by the whole warp (metadata broadcast with shuffles), the first 32
//      ladder entries of row r + 1 in flight while row r is selected; the winners' col offsets go to shared memory
//      instead of being dereferenced;
//   C. the tile's col entries are loaded in one batch (up to f independent loads per lane) and written as one
//      contiguous run of 32 * f outputs.
// Same keys, same (key, idx) order, same outputs as khop_hop_kernel.
constexpr int kTileRows = 32;
#ifndef KHOP_TILE_MIN_BLOCKS
#define KHOP_TILE_MIN_BLOCKS 5  // 46 registers, 40 resident warps per SM: measured 6 / 5 / 4 / 3 blocks -> 0.367 / 0.361 / 0.361 / 0.380 ms per step
#endif

template <int MAXF>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, KHOP_TILE_MIN_BLOCKS) khop_tile_kernel(const HopArgs a) {
    // per warp: the ladder's candidates (truncated keys [160] | hash inputs [160]) or, for rows it does not cover,
    // (keys [64] | positions [64])
    __shared__ uint64_t s_cand[kWarpsPerBlock][kCandCapWide];
    __shared__ int32_t s_off[kWarpsPerBlock][kTileRows * MAXF];  // winner's offset inside its row, -1 = empty, -2 = row deferred
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int f = a.fanouts[a.h - 1];
    const int64_t n_tiles = (a.n_parent + kTileRows - 1) / kTileRows;
    uint32_t* ctk = reinterpret_cast<uint32_t*>(s_cand[w]);
    uint32_t* cx = ctk + kCandCapWide;
    uint64_t* ck = s_cand[w];
    int32_t* ci = reinterpret_cast<int32_t*>(s_cand[w] + kCandCap);
    // persistent warps: tiles are handed out by an atomic counter, so a launch has no tail of half-empty waves and the
    // cheap tiles (empty parents) do not leave SMs idle behind the expensive ones
    for (;;) {
    long long tile = 0;
    if (lane == 0) tile = (long long)atomicAdd(a.tile_counter, 1ULL);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) return;
    const int64_t tile0 = (int64_t)tile * kTileRows;
    const int n_rows = (int)min((int64_t)kTileRows, a.n_parent - tile0);
    const int64_t pslot = tile0 + lane;

    // ---- A: lane-parallel metadata
    int32_t v = -1, mult = 1;
    uint32_t ssum = 0;
    int state = 0;  // 0 = no group (all -1), 1 = select here, 2 = deferred to the heavy pass
    int64_t size = 0, row_begin = 0;
    if (lane < n_rows) {
        bool live = true;
        if (a.h == 1) {
            v = a.roots[pslot];
            ssum = (uint32_t)v;
        } else {
            const int32_t* prev = a.levels[a.h - 2];
            v = prev[pslot];
            live = v >= 0;
            if (live) {
                ssum = (uint32_t)v;
                int64_t s = pslot;
                for (int hh = a.h - 1; hh >= 1; --hh) {
                    s /= a.fanouts[hh - 1];
                    ssum += (uint32_t)(hh == 1 ? a.roots[s] : a.levels[hh - 2][s]);
                }
                const int fp = a.fanouts[a.h - 2];
                const int64_t sib0 = (pslot / fp) * fp;
                int m = 0;
                for (int j = 0; j < fp; ++j) {
                    const bool eq = prev[sib0 + j] == v;
                    m += eq;
                    if (eq && sib0 + j < pslot) live = false;  // a non-first duplicate produces no group
                }
                mult = m;
            }
        }
        if (live) {
            if (v < 0 || v >= a.n_nodes) {
                atomicExch(a.err, GIGL_E_RANGE);
            } else {
                row_begin = __ldg(a.rowptr + v);
                size = (__ldg(a.rowptr + v + 1) - row_begin) * mult;
                if (size > 2147483647LL) {
                    atomicExch(a.err, GIGL_E_OVERFLOW);
                    size = 0;
                } else {
                    state = 1;
                }
            }
        }
    }
    const uint32_t base = ssum + (uint32_t)a.cur_seed;
    const uint32_t size32 = (uint32_t)size;  // state == 1: size < 2^31
    int lvl = -1;
    uint32_t e_beg = 0, e_end = 0, tmax = 0;
    if (state == 1) lad_plan(a, base, size32, f, lvl, e_beg, e_end, tmax);
    if (state == 1 && lvl < 0 && size > kHeavyThreshold && a.heavy_list != nullptr) {
        const int slot = atomicAdd(a.heavy_count, 1);
        if (slot < a.heavy_cap) {
            a.heavy_list[slot] = (int32_t)pslot;
            state = 2;
        }
    }
    const int code = state | ((lvl + 1) << 2);

    // ---- B: one row at a time, whole warp; the next row's first ladder entries are in flight meanwhile
    int32_t* off = s_off[w];
    int n_code = __shfl_sync(0xffffffffu, code, 0);
    uint32_t n_beg = __shfl_sync(0xffffffffu, e_beg, 0), n_end = __shfl_sync(0xffffffffu, e_end, 0);
    uint64_t pre = 0;
    if (n_code >> 2) pre = lad_load(a, (n_code >> 2) - 1, n_beg + lane, n_end);
    for (int r = 0; r < n_rows; ++r) {
        const int r_code = n_code;
        const uint32_t r_beg = n_beg, r_end = n_end;
        const uint64_t cur = pre;
        if (r + 1 < n_rows) {
            n_code = __shfl_sync(0xffffffffu, code, r + 1);
            n_beg = __shfl_sync(0xffffffffu, e_beg, r + 1);
            n_end = __shfl_sync(0xffffffffu, e_end, r + 1);
            if (n_code >> 2) pre = lad_load(a, (n_code >> 2) - 1, n_beg + lane, n_end);
        }
        const int r_state = r_code & 3;
        const uint32_t r_size = __shfl_sync(0xffffffffu, size32, r);
        if (r_state != 1 || r_size == 0) {
            if (lane < f) off[r * f + lane] = (r_state == 2) ? -2 : -1;
            continue;
        }
        const uint32_t r_base = __shfl_sync(0xffffffffu, base, r);
        const int32_t r_mult = __shfl_sync(0xffffffffu, mult, r);
        WarpTopK<1> best;
        best.init();
        bool done = false;
        if (r_code >> 2)
            done = select_ladder(a, best, r_size, r_base, f, lane, (r_code >> 2) - 1, r_beg, r_end, __shfl_sync(0xffffffffu, tmax, r), cur, ctk, cx);
        if (!done && !select_threshold(a, best, (int64_t)r_size, r_base, f, lane, ck, ci)) {
            best.init();
            select_row<1>(a, best, (int64_t)r_size, r_base, f, lane);
        }
        int32_t q = best.idx[0] - 1;
        if (r_mult != 1) q /= r_mult;  // sorted(m copies)[q] = row[q / m]
        if (lane < f) off[r * f + lane] = ((uint32_t)lane < r_size) ? q : -1;
    }
    __syncwarp();

    // ---- C: batched col loads, contiguous writes
    int32_t* out = a.out_nbr + tile0 * f;
    const int total = n_rows * f;
    uint32_t won = 0;  // bit k: this lane's slot of the tile's k-th run of 32 claimed its vertex's stage slot (k < f <= 32)
    for (int i0 = 0; i0 < total; i0 += 32 * 4) {
        int32_t val[4], o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            o[u] = (i < total) ? off[i] : -2;
            const int64_t rb = __shfl_sync(0xffffffffu, row_begin, (i < total) ? i / f : 0);
            val[u] = (o[u] >= 0) ? __ldg(a.col + rb + o[u]) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 32 + lane;
            if (o[u] != -2) out[i] = val[u];  // -2: past the tile, or a row the heavy pass writes
        }
        if (a.st_slot != nullptr) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (stage_try_claim(a, val[u])) won |= 1u << ((i0 >> 5) + u);
        }
    }
    if (a.st_slot != nullptr) {  // one counter atomic per tile: the winners' slots are handed out by a warp scan
        const int mine = __popc(won);
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const int n_won = __shfl_sync(0xffffffffu, incl, 31);
        if (n_won > 0) {  // warp-uniform
            int s0 = 0;
            if (lane == 31) s0 = atomicAdd(a.st_ctr, n_won);
            s0 = __shfl_sync(0xffffffffu, s0, 31) + incl - mine;
            while (won) {
                const int k = __ffs(won) - 1;
                won &= won - 1;
                const int32_t v = out[k * 32 + lane];  // this lane's own write
                a.st_list[s0] = v;
                a.st_slot[v] = s0;
                ++s0;
            }
        }
    }
    if (lane < n_rows && state != 2) a.out_cnt[pslot] = (state == 1) ? (int32_t)min(size, (int64_t)f) : 0;
    __syncwarp();  // s_off is reused by the next tile
    }
}

// One CTA per deferred long row: every warp scans an interleaved share of the window with its own
// top-f list, then warp 0 merges the lists through shared memory.
template <int KPL>
__global__ void __launch_bounds__(kHeavyWarps * 32) khop_heavy_kernel(const HopArgs a) {
    __shared__ uint64_t s_key[kHeavyWarps][32 * KPL];
    __shared__ int32_t s_idx[kHeavyWarps][32 * KPL];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int f = a.fanouts[a.h - 1];
    int n_heavy = *a.heavy_count;
    if (n_heavy > a.heavy_cap) n_heavy = a.heavy_cap;
    for (int w = blockIdx.x; w < n_heavy; w += gridDim.x) {
        const int64_t pslot = a.heavy_list[w];
        int32_t v, mult;
        uint32_t ssum;
        resolve_parent(a, pslot, lane, v, ssum, mult);  // already validated by the first pass
        const int64_t row_begin = __ldg(a.rowptr + v);
        const int64_t size = (__ldg(a.rowptr + v + 1) - row_begin) * mult;
        WarpTopK<KPL> best;
        best.init();
        best.scan((int64_t)warp * 32, size, (int64_t)kHeavyWarps * 32, ssum + (uint32_t)a.cur_seed, f, lane);
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            s_key[warp][k * 32 + lane] = best.key[k];
            s_idx[warp][k * 32 + lane] = best.idx[k];
        }
        __syncthreads();
        if (warp == 0) {
            for (int ow = 1; ow < kHeavyWarps; ++ow) {
                for (int p = 0; p < f; ++p) {
                    const uint64_t ck = s_key[ow][p];
                    if (ck >= best.kth) break;  // lists are sorted ascending
                    best.insert(ck, s_idx[ow][p], f, lane);
                }
            }
            write_result<KPL>(a, pslot, best, size, row_begin, mult, lane);
        }
        __syncthreads();
    }
}


// ---- hash ladder: build -------------------------------------------------------------------------
struct LadderPtrs {
    uint32_t* tk0;
    uint32_t* bs[GIGL_LAD_MAX_LEVELS + 1];   // pass 1: per-block counts shifted by one; after the scans: block starts
    uint32_t* cur[GIGL_LAD_MAX_LEVELS + 1];  // pass 2: fill cursors (a copy of the block starts)
    uint64_t* ent[GIGL_LAD_MAX_LEVELS + 1];
};

// pass 1: level-0 truncated keys and, for every level an input belongs to, its block's count (at index block + 1)
__global__ void __launch_bounds__(256) ladder_count_kernel(uint32_t limit, int levels, const LadderPtrs p) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < limit; x += stride) {
        const uint64_t key = ordered_key((int32_t)x);
        p.tk0[x] = (uint32_t)(key >> 32);
        int z = key ? __clzll((long long)key) : 64;
        if (z > levels) z = levels;
        for (int j = 1; j <= z; ++j) atomicAdd(p.bs[j] + (x >> j) + 1, 1u);
    }
}

// pass 2: the entries, grouped by block (their order inside a block is free: a row's candidates are re-ordered by key)
__global__ void __launch_bounds__(256) ladder_fill_kernel(uint32_t limit, int levels, const LadderPtrs p) {
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t x = blockIdx.x * blockDim.x + threadIdx.x; x < limit; x += stride) {
        const uint64_t key = ordered_key((int32_t)x);
        int z = key ? __clzll((long long)key) : 64;
        if (z > levels) z = levels;
        for (int j = 1; j <= z; ++j) {
            const uint32_t pos = atomicAdd(p.cur[j] + (x >> j), 1u);
            p.ent[j][pos] = ((key >> (32 - j)) << 32) | x;
        }
    }
}

static void ladder_free(gigl_ladder& lad) {
    if (lad.blob_a) cudaFree(lad.blob_a);
    if (lad.blob_b) cudaFree(lad.blob_b);
    lad = gigl_ladder();
}

// Covers hash inputs [0, want).  The ladder is an accelerator, not a requirement: if it cannot be allocated the
// sampler hashes every window on the fly.
static int ensure_ladder(gigl_ctx* ctx, uint64_t want) {
    if (want > (1ULL << 31)) want = 1ULL << 31;
    if (ctx->lad.limit >= want) return GIGL_OK;
    GIGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ladder_free(ctx->lad);
    const uint32_t limit = (uint32_t)want;  // 2^31 fits
    const int levels = GIGL_LAD_MAX_LEVELS;
    auto up = [](size_t v) { return (v + 63) & ~(size_t)63; };
    size_t bs_elems = 0, bs_off[GIGL_LAD_MAX_LEVELS + 1] = {};
    for (int j = 1; j <= levels; ++j) {
        bs_off[j] = bs_elems;
        bs_elems += up((size_t)(limit >> j) + 2);
    }
    const size_t tk_elems = up(limit);
    gigl_ladder lad;
    uint32_t* cursors = nullptr;
    cudaError_t e = cudaMalloc(&lad.blob_a, sizeof(uint32_t) * (tk_elems + bs_elems));
    if (e == cudaSuccess) e = cudaMalloc(&cursors, sizeof(uint32_t) * bs_elems);
    if (e != cudaSuccess) {
        ladder_free(lad);
        if (cursors) cudaFree(cursors);
        cudaGetLastError();
        return GIGL_OK;
    }
    LadderPtrs p{};
    p.tk0 = lad.tk0 = (uint32_t*)lad.blob_a;
    for (int j = 1; j <= levels; ++j) {
        p.bs[j] = lad.bs[j] = lad.tk0 + tk_elems + bs_off[j];
        p.cur[j] = cursors + bs_off[j];
    }
    cudaStream_t st = ctx->stream;
    const int grid = ctx->sm_count * 8;
    GIGL_CUDA(ctx, cudaMemsetAsync(lad.tk0 + tk_elems, 0, sizeof(uint32_t) * bs_elems, st));
    ladder_count_kernel<<<grid, 256, 0, st>>>(limit, levels, p);
    GIGL_LAUNCHED(ctx);
    // counts -> block starts (inclusive sum of the shifted counts; CUB scan = plumbing), one scan per level
    size_t temp_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, temp_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int64_t)(limit >> 1) + 2, st);
    void* temp = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_WORK, temp_bytes + 256, &temp);
    if (rc != GIGL_OK) {
        ladder_free(lad);
        cudaFree(cursors);
        return rc;
    }
    uint32_t totals[GIGL_LAD_MAX_LEVELS + 1] = {};
    for (int j = 1; j <= levels; ++j) {
        const int64_t n = (int64_t)(limit >> j) + 2;
        size_t tb = temp_bytes;
        e = cub::DeviceScan::InclusiveSum(temp, tb, lad.bs[j], lad.bs[j], n, st);
        ctx->launches++;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&totals[j], lad.bs[j] + n - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) break;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(cursors, lad.tk0 + tk_elems, sizeof(uint32_t) * bs_elems, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    size_t ent_elems = 0, ent_off[GIGL_LAD_MAX_LEVELS + 1] = {};
    for (int j = 1; j <= levels; ++j) {
        ent_off[j] = ent_elems;
        ent_elems += up((size_t)totals[j] + 1);
    }
    if (e == cudaSuccess) e = cudaMalloc(&lad.blob_b, sizeof(uint64_t) * ent_elems);
    if (e != cudaSuccess) {
        ladder_free(lad);
        cudaFree(cursors);
        cudaGetLastError();
        return GIGL_OK;
    }
    for (int j = 1; j <= levels; ++j) p.ent[j] = lad.ent[j] = (uint64_t*)lad.blob_b + ent_off[j];
    ladder_fill_kernel<<<grid, 256, 0, st>>>(limit, levels, p);
    GIGL_LAUNCHED(ctx);
    GIGL_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(cursors);
    lad.limit = limit;
    lad.levels = levels;
    ctx->lad = lad;
    return GIGL_OK;
}

template <int KPL>
static int launch_hop(gigl_ctx* ctx, const HopArgs& a, bool tiled) {
    const int64_t blocks = ceil_div64(a.n_parent, kWarpsPerBlock * (tiled && KPL == 1 ? kTileRows : 1));
    if (blocks > 0x7fffffffLL) return gigl_fail(ctx, GIGL_E_INVALID, "too many frontier slots for one launch");
    if (blocks > 0) {
        if (tiled && KPL == 1) {
            const int64_t cap = (int64_t)ctx->sm_count * KHOP_TILE_MIN_BLOCKS;
            const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
            if (a.fanouts[a.h - 1] <= 16)
                khop_tile_kernel<16><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a);
            else
                khop_tile_kernel<32><<<grid, kWarpsPerBlock * 32, 0, ctx->stream>>>(a);
        } else {
            khop_hop_kernel<KPL><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, ctx->stream>>>(a);
        }
        GIGL_LAUNCHED(ctx);
        if (a.heavy_list != nullptr) {
            const int grid = ctx->sm_count * 2;
            khop_heavy_kernel<KPL><<<grid, kHeavyWarps * 32, 0, ctx->stream>>>(a);
            GIGL_LAUNCHED(ctx);
        }
    }
    return GIGL_OK;
}

}  // namespace gigl

int khop_sample_launch(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                       int32_t n_hops, int32_t base_seed, int32_t first_call_no, int32_t* const* nbr_dev,
                       int32_t* const* cnt_dev, int32_t hop_first, int32_t hop_last, const gigl_stage_args* stage,
                       const float* weights_dev, int32_t method) {
    using namespace gigl;
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, n_hops >= 1 && n_hops <= GIGL_MAX_HOPS, "n_hops must be in [1, 8]");
    GIGL_CHECK(ctx, n_roots >= 0, "n_roots < 0");
    GIGL_CHECK(ctx, fanouts && nbr_dev && cnt_dev, "null fanouts / output tables");
    GIGL_CHECK(ctx, roots_dev || n_roots == 0, "null roots");
    GIGL_CHECK(ctx, method == GIGL_SAMPLE_UNIFORM || ((method == GIGL_SAMPLE_TOP_K || method == GIGL_SAMPLE_RANDOM_WEIGHTED) && weights_dev),
               "sampling method must be uniform, or top-k / random-weighted with an edge-weight array");
    for (int h = 0; h < n_hops; ++h) {
        GIGL_CHECK(ctx, fanouts[h] >= 1 && fanouts[h] <= GIGL_MAX_FANOUT, "fanout must be in [1, 128]");
        GIGL_CHECK(ctx, (nbr_dev[h] && cnt_dev[h]) || n_roots == 0, "null output level");
    }
    if (n_roots == 0) return GIGL_OK;

    // heavy-row worklist: one int32 counter + up to heavy_cap slots, reused hop after hop
    const int32_t heavy_cap = 1 << 22;
    int rc;
    // hash inputs of a window: path-id sum (< n_hops * n_nodes) + seed * call number + position; the margin covers the
    // seed term and the row length of all but the longest rows' tails
    if (g->hx_enabled && (rc = ensure_ladder(ctx, (uint64_t)n_hops * (uint64_t)g->n_nodes + (1ULL << 21))) != GIGL_OK) return rc;
    void* scratch = nullptr;
    rc = gigl_scratch(ctx, GIGL_SLOT_WORK, sizeof(int32_t) * ((size_t)heavy_cap + 64), &scratch);
    if (rc != GIGL_OK) return rc;
    int32_t* heavy_count = (int32_t*)scratch;
    int32_t* heavy_list = heavy_count + 64;

    HopArgs a{};
    if (g->hx_enabled && ctx->lad.limit) {
        a.lad_tk0 = ctx->lad.tk0;
        for (int j = 1; j <= ctx->lad.levels; ++j) {
            a.lad_bs[j] = ctx->lad.bs[j];
            a.lad_ent[j] = ctx->lad.ent[j];
        }
        a.lad_limit = ctx->lad.limit;
        a.lad_levels = ctx->lad.levels;
    }
    a.rowptr = g->rowptr;
    a.col = g->col;
    a.n_nodes = g->n_nodes;
    a.roots = roots_dev;
    a.err = ctx->d_err;
    a.heavy_list = heavy_list;
    a.heavy_count = heavy_count;
    a.heavy_cap = heavy_cap;
    a.tile_counter = (unsigned long long*)(heavy_count + 2);
    if (stage != nullptr) {
        a.st_slot = stage->slot;
        a.st_list = stage->list;
        a.st_ctr = stage->ctr;
    }
    int64_t n_parent = n_roots;
    static const bool tiled = [] {  // GIGL_KHOP_TILE=0: the warp-per-row kernel (A/B measurements)
        const char* e = getenv("GIGL_KHOP_TILE");
        return !(e && e[0] == '0');
    }();
    gigl_timed timed(ctx, GIGL_T_SAMPLE);
    if (hop_last <= 0 || hop_last > n_hops) hop_last = n_hops;
    for (int h = 1; h <= hop_last; ++h) {
        const int32_t f = fanouts[h - 1];
        a.fanouts[h - 1] = f;
        if (h >= 2) a.levels[h - 2] = nbr_dev[h - 2];
        if (h < hop_first) {  // levels below hop_first were sampled by an earlier call
            n_parent *= f;
            continue;
        }
        a.h = h;
        a.cur_seed = (int32_t)((uint32_t)base_seed * ((uint32_t)first_call_no + (uint32_t)(h - 1)));
        a.out_nbr = nbr_dev[h - 1];
        a.out_cnt = cnt_dev[h - 1];
        a.n_parent = n_parent;
        if (n_parent > 0x7fffffffLL) return gigl_fail(ctx, GIGL_E_INVALID, "frontier exceeds 2^31-1 slots; split the roots");
        GIGL_CUDA(ctx, cudaMemsetAsync(heavy_count, 0, sizeof(int32_t) * 4, ctx->stream));  // + the tile counter
        if (method != GIGL_SAMPLE_UNIFORM) {
            const int64_t blocks = ceil_div64(n_parent, kWarpsPerBlock);
            if (blocks > 0x7fffffffLL) return gigl_fail(ctx, GIGL_E_INVALID, "too many frontier slots for one launch");
            if (blocks > 0) {
                if (f <= 32)
                    khop_weighted_kernel<1><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, weights_dev, method);
                else if (f <= 64)
                    khop_weighted_kernel<2><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, weights_dev, method);
                else
                    khop_weighted_kernel<4><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, ctx->stream>>>(a, weights_dev, method);
                GIGL_LAUNCHED(ctx);
            }
            rc = GIGL_OK;
        } else if (f <= 32)
            rc = launch_hop<1>(ctx, a, tiled);
        else if (f <= 64)
            rc = launch_hop<2>(ctx, a, false);
        else
            rc = launch_hop<4>(ctx, a, false);
        if (rc != GIGL_OK) return rc;
        n_parent *= f;
    }
    return GIGL_OK;
}
