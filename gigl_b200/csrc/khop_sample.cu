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
#include <stdlib.h>

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
    // hash-window index (see build_hash_index_kernel): per block of 2^l_log2 consecutive hash inputs
    // the `cap` smallest keys, ascending, with their offsets inside the block; nullptr = disabled
    const uint64_t* hx_keys;
    const uint16_t* hx_offs;
    uint64_t hx_limit;  // inputs [0, hx_limit) are covered
    int32_t hx_l_log2;
    int32_t hx_cap;
    // full key table: hk_table[x] = ordered_key(x) for x in [0, hk_limit); nullptr = hash on the fly
    const uint64_t* hk_table;
    uint64_t hk_limit;
};

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
        if (p < f) {
            int32_t r = -1;
            if (p < n_out) {
                const int64_t j = (int64_t)(best.idx[k] - 1) / mult;  // sorted(m copies)[q] = row[q / m]
                r = __ldg(a.col + row_begin + j);
            }
            o[p] = r;
        }
    }
    if (lane == 0) a.out_cnt[pslot] = n_out;
}

// ---- hash-window index ----------------------------------------------------------------------
// The permutation key of window position i is H(base + i): a row of `size` entries asks for the f
// smallest values of the FIXED sequence H(x) over the window x in [base+1, base+size].  For long
// rows (hubs) almost all of that hashing is shared between windows, so the sequence is indexed
// once per graph: for every aligned block of L = 2^l_log2 inputs the `cap` smallest keys (sorted)
// and their offsets.  A window then costs < 2L hashes (its unaligned head and tail) plus one
// 8-byte read per covered block (the block minimum prunes nearly all of them), instead of `size`
// hashes.  Bit-exact by construction: the same keys, the same (key, idx) order.
template <int KPL>
__global__ void __launch_bounds__(256) build_hash_index_kernel(int64_t n_blocks, int l_log2, int cap,
                                                               uint64_t* __restrict__ keys,
                                                               uint16_t* __restrict__ offs) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n_blocks; b += warps) {
        WarpTopK<KPL> best;
        best.init();
        const uint32_t base = (uint32_t)((b << l_log2) - 1);  // x = base + i, i = 1..L
        best.scan(0, (int64_t)1 << l_log2, 32, base, cap, lane);
#pragma unroll
        for (int k = 0; k < KPL; ++k) {
            const int p = k * 32 + lane;
            if (p < cap) {
                keys[b * cap + p] = best.key[k];
                offs[b * cap + p] = (uint16_t)(best.idx[k] - 1);
            }
        }
    }
}

// Top-f of one row's window, through the index when it covers the window.
template <int KPL>
__device__ __forceinline__ void select_row(const HopArgs& a, WarpTopK<KPL>& best, int64_t size, uint32_t base, int f,
                                           int lane) {
    if (a.hx_keys != nullptr) {
        const uint64_t lo = (uint64_t)base + 1, hi = (uint64_t)base + (uint64_t)size;
        const int lg = a.hx_l_log2;
        if (hi < a.hx_limit && (size >> lg) != 0) {
            const uint64_t b0 = (lo + ((1ULL << lg) - 1)) >> lg, b1 = (hi + 1) >> lg;  // full blocks [b0, b1)
            if (b0 < b1) {
                const int64_t head_end = (int64_t)((b0 << lg) - lo);             // i in [1, head_end]
                const int64_t tail_begin = (int64_t)((b1 << lg) - (uint64_t)base);  // first tail i
                best.scan(0, head_end, 32, base, f, lane);
                best.scan(tail_begin - 1, size, 32, base, f, lane);
                const int cap = a.hx_cap;
                for (uint64_t bb = b0; bb < b1; bb += 32) {
                    const uint64_t myb = bb + lane;
                    uint64_t head = kKeyInf;
                    if (myb < b1) head = __ldg(a.hx_keys + myb * cap);
                    uint32_t cand = __ballot_sync(0xffffffffu, head < best.kth);
                    while (cand) {
                        const int src = __ffs(cand) - 1;
                        cand &= cand - 1;
                        if (__shfl_sync(0xffffffffu, head, src) >= best.kth) continue;
                        const uint64_t blk = bb + src;
#pragma unroll
                        for (int k = 0; k < KPL; ++k) {
                            const int p = k * 32 + lane;
                            uint64_t ek = kKeyInf;
                            uint32_t eo = 0;
                            if (p < cap) {
                                ek = __ldg(a.hx_keys + blk * cap + p);
                                eo = __ldg(a.hx_offs + blk * cap + p);
                            }
                            uint32_t c2 = __ballot_sync(0xffffffffu, ek < best.kth);
                            while (c2) {
                                const int s2 = __ffs(c2) - 1;
                                c2 &= c2 - 1;
                                const uint64_t ck = __shfl_sync(0xffffffffu, ek, s2);
                                const uint32_t co = __shfl_sync(0xffffffffu, eo, s2);
                                if (ck >= best.kth) break;  // entries are ascending
                                const uint32_t x = (uint32_t)(blk << lg) + co;
                                best.insert(ck, (int32_t)(x - base), f, lane);
                            }
                        }
                    }
                }
                return;
            }
        }
    }
    best.scan(0, size, 32, base, f, lane);
}

// ---- threshold selection (fanout <= 32) ------------------------------------------------------
// The keys are uniform 64-bit values, so the f-th smallest of `size` keys sits near (f / size) * 2^64.
// One pass collects every key below T = (m / size) * 2^64, m a little above f, into a per-warp
// shared-memory buffer (no serial insertions), one bitonic sort orders the <= 64 candidates, the
// first f are the answer.  If fewer than f keys fell below T, or more than the buffer holds (both
// rare by the Poisson tail), the exact streaming path above redoes the row - the result is always
// the exact top-f of the exact keys.
constexpr int kCandCap = 64;

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

__device__ __forceinline__ bool scan_threshold(const HopArgs& a, bool use_tab, uint32_t base, int64_t i_first, int64_t i_last,
                                               uint64_t T, uint64_t* ck, int32_t* ci, int& c, int lane) {
    for (int64_t c0 = i_first - 1; c0 < i_last; c0 += 32) {
        const int64_t i = c0 + lane + 1;
        uint64_t k = kKeyInf;
        if (i <= i_last) {
            const uint32_t x = base + (uint32_t)i;
            k = use_tab ? __ldg(a.hk_table + x) : ordered_key((int32_t)x);
        }
        if (!push_cands(ck, ci, c, k < T, k, (int32_t)i, lane)) return false;
    }
    return true;
}

// The same scan over hash inputs x in [x_first, x_last], all below 2^31 (index-covered rows): 32-bit arithmetic only.
__device__ __forceinline__ bool scan_threshold32(const HopArgs& a, bool use_tab, uint32_t base, uint32_t x_first, uint32_t x_last,
                                                 uint64_t T, uint64_t* ck, int32_t* ci, int& c, int lane) {
    for (uint32_t x0 = x_first; x0 <= x_last; x0 += 32) {
        const uint32_t x = x0 + (uint32_t)lane;
        uint64_t k = kKeyInf;
        if (x <= x_last) k = use_tab ? __ldg(a.hk_table + x) : ordered_key((int32_t)x);
        if (!push_cands(ck, ci, c, k < T, k, (int32_t)(x - base), lane)) return false;
    }
    return true;
}

// Returns true with best.key[0] / best.idx[0] = the row's keys in ascending order (>= f of them, or all).
__device__ __forceinline__ bool select_threshold(const HopArgs& a, WarpTopK<1>& best, int64_t size, uint32_t base, int f,
                                                 int lane, uint64_t* ck, int32_t* ci) {
    float m = (float)f + fmaxf(14.f, 1.2f * (float)f);
    if (m > 48.f) m = 48.f;
    uint64_t T = kKeyInf;
    if ((float)size > m) T = __float2ull_rz(m / (float)size * 18446744073709551616.0f);
    const uint64_t hi = (uint64_t)base + (uint64_t)size;
    const bool use_tab = a.hk_table != nullptr && hi < a.hk_limit;
    int c = 0;
    bool done = false;
    if (a.hx_keys != nullptr && hi < a.hx_limit) {
        // Every hash input of the window lies below hx_limit <= 2^31: the window, its blocks and the candidates'
        // positions are 32-bit quantities (ncu: 52 % of this kernel's issue slots were 64-bit index arithmetic).
        const uint32_t lo32 = base + 1u, hi32 = (uint32_t)hi;
        const int lg = a.hx_l_log2;
        const uint32_t b0 = (lo32 + ((1u << lg) - 1u)) >> lg, b1 = (hi32 + 1u) >> lg;  // full blocks [b0, b1)
        if (b0 < b1) {
            if (!scan_threshold32(a, use_tab, base, lo32, (b0 << lg) - 1u, T, ck, ci, c, lane)) return false;   // head
            if (!scan_threshold32(a, use_tab, base, b1 << lg, hi32, T, ck, ci, c, lane)) return false;          // tail
            const uint32_t cap = (uint32_t)a.hx_cap;
            for (uint32_t bb = b0; bb < b1; bb += 32) {
                const uint32_t myb = bb + (uint32_t)lane;
                bool active = myb < b1;
                const uint64_t* bk = a.hx_keys + (size_t)myb * cap;
                const uint16_t* bo = a.hx_offs + (size_t)myb * cap;
                for (uint32_t j = 0; __any_sync(0xffffffffu, active); ++j) {
                    uint64_t ek = kKeyInf;
                    if (active) ek = __ldg(bk + j);
                    const bool cand = active && ek < T;
                    uint32_t x = 0;
                    if (cand) x = (myb << lg) + (uint32_t)__ldg(bo + j);  // offsets are read for candidates only
                    if (!push_cands(ck, ci, c, cand, ek, (int32_t)(x - base), lane)) return false;
                    active = cand && (j + 1 < cap);
                }
            }
        } else if (!scan_threshold32(a, use_tab, base, lo32, hi32, T, ck, ci, c, lane)) {
            return false;
        }
        done = true;
    }
    if (!done && !scan_threshold(a, use_tab, base, 1, size, T, ck, ci, c, lane)) return false;
    const int64_t need = size < f ? size : f;
    if (c < need) return false;
    __syncwarp();
    if (need < 32) {
        // Order the candidates on ONE 32-bit word each: the top 26 significant bits of the key (every candidate is
        // below T, so bits above T's highest bit are zero) over the candidate's 6-bit buffer slot.  A 32-lane bitonic
        // network on such words is one shuffle and one min/max per stage instead of three shuffles and a 64-bit
        // compare.  Two candidates agreeing on those 26 bits inside the first need + 1 positions (about 1e-5 of the
        // rows) would make the order ambiguous: the row is then redone by the exact streaming path.
        int shift = 38;
        if (T != kKeyInf) {
            const int nbits = 64 - __clzll((long long)T);
            shift = nbits > 26 ? nbits - 26 : 0;
        }
        uint32_t v0 = 0xFFFFFFFFu, v1 = 0xFFFFFFFFu;
        if (lane < c) v0 = ((uint32_t)(ck[lane] >> shift) << 6) | (uint32_t)lane;
        sort32_u32(v0, lane);
        if (c > 32) {  // warp-uniform
            if (lane + 32 < c) v1 = ((uint32_t)(ck[lane + 32] >> shift) << 6) | (uint32_t)(lane + 32);
            sort32_u32(v1, lane);
            // 32 smallest of two ascending runs: min(v0[l], v1[31 - l]) is bitonic; one merge sorts it
            v0 = min(v0, __shfl_sync(0xffffffffu, v1, 31 - lane));
#pragma unroll
            for (int stride = 16; stride > 0; stride >>= 1) cmpx_u32(v0, stride, (lane & stride) == 0);
        }
        const uint32_t up = __shfl_up_sync(0xffffffffu, v0, 1);
        const bool tie = lane >= 1 && lane <= need && v0 != 0xFFFFFFFFu && (v0 >> 6) == (up >> 6);
        if (__any_sync(0xffffffffu, tie)) return false;
        if (lane < need) {
            best.key[0] = ck[v0 & 63u];
            best.idx[0] = ci[v0 & 63u];
        }
        __syncwarp();
        return true;
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
        // 32 smallest of two ascending runs: min(v0[l], v1[31 - l]) is bitonic; one merge sorts it
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

__device__ __forceinline__ bool row_uses_index(const HopArgs& a, int64_t size, uint32_t base) {
    if (a.hx_keys == nullptr) return false;
    const uint64_t lo = (uint64_t)base + 1, hi = (uint64_t)base + (uint64_t)size;
    const int lg = a.hx_l_log2;
    return hi < a.hx_limit && ((lo + ((1ULL << lg) - 1)) >> lg) < ((hi + 1) >> lg);
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
    if (size > kHeavyThreshold && a.heavy_list != nullptr && !row_uses_index(a, size, base)) {
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
        __shared__ uint64_t s_ck[kWarpsPerBlock][kCandCap];
        __shared__ int32_t s_ci[kWarpsPerBlock][kCandCap];
        const int w = threadIdx.x >> 5;
        if (size <= 32 || !select_threshold(a, best, size, base, f, lane, s_ck[w], s_ci[w])) {
            best.init();
            select_row<KPL>(a, best, size, base, f, lane);
        }
    } else {
        select_row<KPL>(a, best, size, base, f, lane);
    }
    write_result<KPL>(a, pslot, best, size, row_begin, mult, lane);
}

// ---- tile form of the hop kernel (fanout <= 32) ---------------------------------------------------
// khop_hop_kernel is latency-bound (ncu: 41 % of the stall samples wait on global loads, DRAM 7 % busy): every warp
// walks ONE row through four dependent round trips (parent slot -> rowptr pair -> key window -> winning col entries).
// Here a warp owns a tile of 32 consecutive parent slots:
//   A. lane l resolves slot l's metadata (parent vertex, path-id sum, sibling multiplicity, rowptr pair): the same
//      loads, but 32 rows share each round trip;
//   B. the rows are selected one after the other by the whole warp (metadata broadcast with shuffles); the winners'
//      col offsets go to shared memory instead of being dereferenced;
//   C. the tile's col entries are loaded in one batch (up to f independent loads per lane) and written as one
//      contiguous run of 32 * f outputs.
// Same keys, same (key, idx) order, same outputs as khop_hop_kernel.
constexpr int kTileRows = 32;
#ifndef KHOP_TILE_MIN_BLOCKS
#define KHOP_TILE_MIN_BLOCKS 6
#endif

template <int MAXF>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, KHOP_TILE_MIN_BLOCKS) khop_tile_kernel(const HopArgs a) {
    __shared__ uint64_t s_ck[kWarpsPerBlock][kCandCap];
    __shared__ int32_t s_ci[kWarpsPerBlock][kCandCap];
    __shared__ int32_t s_off[kWarpsPerBlock][kTileRows * MAXF];  // winner's offset inside its row, -1 = empty, -2 = row deferred
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int f = a.fanouts[a.h - 1];
    const int64_t n_tiles = (a.n_parent + kTileRows - 1) / kTileRows;
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
    if (state == 1 && size > kHeavyThreshold && a.heavy_list != nullptr && !row_uses_index(a, size, base)) {
        const int slot = atomicAdd(a.heavy_count, 1);
        if (slot < a.heavy_cap) {
            a.heavy_list[slot] = (int32_t)pslot;
            state = 2;
        }
    }

    // ---- B: one row at a time, whole warp
    int32_t* off = s_off[w];
    for (int r = 0; r < n_rows; ++r) {
        const int r_state = __shfl_sync(0xffffffffu, state, r);
        const int64_t r_size = __shfl_sync(0xffffffffu, size, r);
        if (r_state != 1 || r_size == 0) {
            if (lane < f) off[r * f + lane] = (r_state == 2) ? -2 : -1;
            continue;
        }
        const uint32_t r_base = __shfl_sync(0xffffffffu, base, r);
        const int32_t r_mult = __shfl_sync(0xffffffffu, mult, r);
        WarpTopK<1> best;
        best.init();
        if (!select_threshold(a, best, r_size, r_base, f, lane, s_ck[w], s_ci[w])) {
            best.init();
            select_row<1>(a, best, r_size, r_base, f, lane);
        }
        int32_t q = best.idx[0] - 1;
        if (r_mult != 1) q /= r_mult;  // sorted(m copies)[q] = row[q / m]
        if (lane < f) off[r * f + lane] = (lane < r_size) ? q : -1;
    }
    __syncwarp();

    // ---- C: batched col loads, contiguous writes
    int32_t* out = a.out_nbr + tile0 * f;
    const int total = n_rows * f;
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

__global__ void build_key_table_kernel(int64_t n, uint64_t* __restrict__ table) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) table[x] = ordered_key((int32_t)(uint32_t)x);
}

static int ensure_hash_index(gigl_graph* g, int fmax, int n_hops) {
    gigl_ctx* ctx = g->ctx;
    if (!g->hx_enabled) return GIGL_OK;
    const int cap = fmax <= 16 ? 16 : fmax <= 32 ? 32 : fmax <= 64 ? 64 : 128;
    int lg = 0;
    static const int lmul = getenv("GIGL_HX_LMUL") ? atoi(getenv("GIGL_HX_LMUL")) : 8;  // block length / cap (experiments)
    while ((1 << lg) < lmul * cap) ++lg;
    uint64_t want = (uint64_t)n_hops * (uint64_t)g->n_nodes + (1ULL << 21);
    if (want > (1ULL << 31)) want = 1ULL << 31;
    const int64_t n_blocks = (int64_t)(want >> lg);
    const uint64_t limit = (uint64_t)n_blocks << lg;
    if (g->hx_keys && g->hx_cap == cap && g->hx_limit >= limit) return GIGL_OK;
    GIGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (g->hx_keys) cudaFree(g->hx_keys);
    if (g->hx_offs) cudaFree(g->hx_offs);
    g->hx_keys = nullptr;
    g->hx_offs = nullptr;
    g->hx_limit = 0;
    if (n_blocks == 0) return GIGL_OK;
    cudaError_t e = cudaMalloc(&g->hx_keys, sizeof(uint64_t) * (size_t)n_blocks * cap);
    if (e == cudaSuccess) e = cudaMalloc(&g->hx_offs, sizeof(uint16_t) * (size_t)n_blocks * cap);
    if (e != cudaSuccess) {  // the index is an accelerator, not a requirement: run without it
        if (g->hx_keys) cudaFree(g->hx_keys);
        g->hx_keys = nullptr;
        g->hx_offs = nullptr;
        cudaGetLastError();
        return GIGL_OK;
    }
    const int grid = ctx->sm_count * 8;
    if (cap <= 32)
        build_hash_index_kernel<1><<<grid, 256, 0, ctx->stream>>>(n_blocks, lg, cap, g->hx_keys, g->hx_offs);
    else if (cap == 64)
        build_hash_index_kernel<2><<<grid, 256, 0, ctx->stream>>>(n_blocks, lg, cap, g->hx_keys, g->hx_offs);
    else
        build_hash_index_kernel<4><<<grid, 256, 0, ctx->stream>>>(n_blocks, lg, cap, g->hx_keys, g->hx_offs);
    GIGL_LAUNCHED(ctx);
    g->hx_cap = cap;
    g->hx_l_log2 = lg;
    g->hx_limit = limit;
    // the full key table (8 bytes per hash input): short windows read their keys instead of hashing them
    if (g->hk_table) cudaFree(g->hk_table);
    g->hk_table = nullptr;
    g->hk_limit = 0;
    if (cudaMalloc(&g->hk_table, sizeof(uint64_t) * (size_t)limit) == cudaSuccess) {
        build_key_table_kernel<<<grid, 256, 0, ctx->stream>>>((int64_t)limit, g->hk_table);
        GIGL_LAUNCHED(ctx);
        g->hk_limit = limit;
    } else {
        g->hk_table = nullptr;
        cudaGetLastError();
    }
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
                       int32_t* const* cnt_dev, int32_t hop_first, int32_t hop_last) {
    using namespace gigl;
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, n_hops >= 1 && n_hops <= GIGL_MAX_HOPS, "n_hops must be in [1, 8]");
    GIGL_CHECK(ctx, n_roots >= 0, "n_roots < 0");
    GIGL_CHECK(ctx, fanouts && nbr_dev && cnt_dev, "null fanouts / output tables");
    GIGL_CHECK(ctx, roots_dev || n_roots == 0, "null roots");
    for (int h = 0; h < n_hops; ++h) {
        GIGL_CHECK(ctx, fanouts[h] >= 1 && fanouts[h] <= GIGL_MAX_FANOUT, "fanout must be in [1, 128]");
        GIGL_CHECK(ctx, (nbr_dev[h] && cnt_dev[h]) || n_roots == 0, "null output level");
    }
    if (n_roots == 0) return GIGL_OK;

    // heavy-row worklist: one int32 counter + up to heavy_cap slots, reused hop after hop
    const int32_t heavy_cap = 1 << 22;
    void* scratch = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_WORK, sizeof(int32_t) * ((size_t)heavy_cap + 64), &scratch);
    if (rc != GIGL_OK) return rc;
    int32_t* heavy_count = (int32_t*)scratch;
    int32_t* heavy_list = heavy_count + 64;

    int fmax = 0;
    for (int h = 0; h < n_hops; ++h) fmax = fanouts[h] > fmax ? fanouts[h] : fmax;
    if ((rc = ensure_hash_index(g, fmax, n_hops)) != GIGL_OK) return rc;

    HopArgs a{};
    a.hx_keys = g->hx_enabled ? g->hx_keys : nullptr;
    a.hx_offs = g->hx_offs;
    a.hx_limit = g->hx_limit;
    a.hx_l_log2 = g->hx_l_log2;
    a.hx_cap = g->hx_cap;
    a.hk_table = (g->hx_enabled && a.hx_keys) ? g->hk_table : nullptr;
    a.hk_limit = g->hk_limit;
    a.rowptr = g->rowptr;
    a.col = g->col;
    a.n_nodes = g->n_nodes;
    a.roots = roots_dev;
    a.err = ctx->d_err;
    a.heavy_list = heavy_list;
    a.heavy_count = heavy_count;
    a.heavy_cap = heavy_cap;
    a.tile_counter = (unsigned long long*)(heavy_count + 2);
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
        if (f <= 32)
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
