// batch_collate.cu - B sampled rooted neighbourhoods -> ONE coalesced batch graph, on the device,
// and the layer-wise SAGE forward over it.
//
// Replaces the Python batch construction that defines what `x` / `edge_index` mean for the
// aggregate:
//   GbmlProtosTranslator.graph_data_from_GraphPb  python/gigl/src/common/translators/gbml_protos_translator.py:101-121
//   GraphBuilder.add_graph_data / PygGraphBuilder.build   python/gigl/src/common/graph_builder/abstract_graph_builder.py:49-197,
//                                                          pyg_graph_builder.py:20-69
//   collate fns + coalesce   python/gigl/src/training/v1/lib/data_loaders/rooted_node_neighborhood_data_loader.py:78-,
//                            supervised_node_classification_data_loader.py:73-117, data_loaders/utils.py:59-146
// i.e. the union of the B subgraphs with nodes de-duplicated by id and edges de-duplicated by
// (src, dst); the model is then run on that union and the root rows are selected
// (graphsage_template_modeling_spec.py:305-311, :565-577).
//
// Device formulation (no per-node Python dicts, no x_batch copy):
//   1. every filled tree slot becomes a 64-bit key (dst << 32 | src); one radix sort (CUB, plumbing)
//      orders the batch's edges by (dst, src): equal keys = duplicate edges, runs of equal dst =
//      the in-edge row of that vertex in the coalesced graph;
//   2. a dense per-vertex map (HBM is large: 8 B x N) records each row's [beg, end) in the sorted keys;
//   3. local ids are assigned level by level: S_L = roots, S_{l-1} = S_l + in-neighbours(S_l), so
//      layer l only computes the rows the root outputs depend on (identical root embeddings to
//      running every layer on every batch node, which is what the reference does);
//   4. the gather kernel walks a row's sorted keys, skips duplicates, pulls the source feature rows
//      (layer 1: straight from the graph-wide feature table by global id; deeper layers: through
//      the local-id map) and writes [mean | self] rows that the projection GEMM consumes.
#include <cuda_runtime.h>
#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "common.cuh"
#include "tcgen05.cuh"

namespace gigl {

constexpr int32_t kLidAbsent = 0x7fffffff;
constexpr int32_t kLidPending = 0x7ffffffe;

static inline int bits_for64(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

// ---- 1. tree slots -> edge keys --------------------------------------------------------------
// key = dst << shift | src with shift = bits(n_graph_nodes): only the used bits are sorted.  Filled
// slots are written densely (warp-aggregated cursor); their order before the sort is irrelevant.
constexpr int kKeysPerThread = 4;
__global__ void __launch_bounds__(256) tree_keys_kernel(int64_t n_slots, int32_t f, const int32_t* __restrict__ parents,
                                                        const int32_t* __restrict__ children, int shift,
                                                        uint64_t* __restrict__ keys, unsigned long long* __restrict__ cursor) {
    __shared__ int s_warp[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per_block = 256 * kKeysPerThread;
    const int64_t n_iter = (n_slots + per_block - 1) / per_block;
    for (int64_t it = blockIdx.x; it < n_iter; it += gridDim.x) {  // block-uniform trip count
        // thread t owns slots base + t*4 .. +3 (contiguous per thread: 16-byte loads)
        const int64_t s0 = it * per_block + (int64_t)threadIdx.x * kKeysPerThread;
        int32_t src[kKeysPerThread];
        int mine = 0;
#pragma unroll
        for (int q = 0; q < kKeysPerThread; ++q) {
            src[q] = (s0 + q < n_slots) ? __ldg(children + s0 + q) : -1;
            mine += src[q] >= 0;
        }
        // exclusive prefix of `mine` inside the warp, then across the 8 warps, then ONE global atomic per block
        int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int t = s_warp[w];
                s_warp[w] = tot;
                tot += t;
            }
            s_base = tot ? atomicAdd(cursor, (unsigned long long)tot) : 0ULL;
        }
        __syncthreads();
        unsigned long long pos = s_base + (unsigned long long)(s_warp[warp] + incl - mine);
#pragma unroll
        for (int q = 0; q < kKeysPerThread; ++q) {
            if (src[q] >= 0) {
                const int32_t dst = __ldg(parents + (s0 + q) / f);
                keys[pos++] = ((uint64_t)(uint32_t)dst << shift) | (uint32_t)src[q];
            }
        }
        __syncthreads();
    }
}

// ---- 2. row boundaries in the sorted keys -----------------------------------------------------
// counters[1] = number of unique edges
__global__ void seg_bounds_kernel(int64_t n_keys, const uint64_t* __restrict__ keys, int shift,
                                  int2* __restrict__ segmap, int32_t* __restrict__ counters) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int uniq = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys; i += stride) {
        const uint64_t k = keys[i];
        const uint32_t d = (uint32_t)(k >> shift);
        const uint64_t kp = (i > 0) ? keys[i - 1] : ~k;
        if (i == 0 || (uint32_t)(kp >> shift) != d) segmap[d].x = (int32_t)i;
        if (i + 1 == n_keys || (uint32_t)(keys[i + 1] >> shift) != d) segmap[d].y = (int32_t)(i + 1);
        uniq += (kp != k);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) uniq += __shfl_xor_sync(0xffffffffu, uniq, off);
    if ((threadIdx.x & 31) == 0 && uniq) atomicAdd(counters + 1, uniq);
}

__global__ void seg_clear_kernel(int64_t n_keys, const uint64_t* __restrict__ keys, int shift,
                                 int2* __restrict__ segmap) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_keys; i += stride) {
        const uint64_t k = keys[i];
        if (i == 0 || (keys[i - 1] >> shift) != (k >> shift)) segmap[(uint32_t)(k >> shift)] = make_int2(0, 0);
    }
}

// ---- 1' / 2'. bucketed form: edges grouped by destination row first, every row sorted on its own ------------------
// One global radix sort of (dst, src) keys moves every key six times (0.48 ms of a 2.5 ms step).  Grouping by dst is a
// counting sort on a dense per-vertex array (HBM is large), and what remains are 375 k INDEPENDENT sorts of 17 sources
// on average - a warp's registers for rows of <= 512, a CTA's shared memory beyond, value-range partitions for hubs:
//   count:    one atomicAdd(segmap[dst].x, children) per (warp, parent slot); the first toucher lists the row;
//   alloc:    block scan of the listed rows' counts, one cursor atomic per block -> row [beg, end) in the key array;
//   scatter:  the same walk, cursor atomic on segmap[dst].y, 4-byte sources into the row's bucket;
//   sort:     per row, ascending by source; the 64-bit keys (dst << shift | src) are written once, in place.
// Duplicate edges stay adjacent in the sorted row (consumers skip equal neighbours), exactly what the global sort gave.
// Counters in d_ctr: [0] listed rows, [1] unique edges, [2] node counter, [3] valid keys, [4 + j] level ends, [16] long
// items, [17] / [18] work-queue cursors of the sort launch.
constexpr int kCtrRows = 0, kCtrUnique = 1, kCtrNodes = 2, kCtrValid = 3, kCtrLong = 16;
constexpr int kWarpSortMax = 512;   // rows up to this many entries are sorted in one warp's registers
constexpr int kLongCap = 4096;      // shared-memory sort capacity of a long-row CTA
constexpr int kLongPart = 1024;     // expected entries per value-range partition of a row longer than kLongCap
constexpr int kLongThreads = 512;

// Slot-parallel walk of one tree level, kRowsUnroll x 32 consecutive slots per warp and iteration (the loads of all
// of them are issued before the first atomic: the walk is bound by round trips, not by bytes).  The lanes of one parent
// slot inside a warp form a segment; its lowest lane speaks for it.
constexpr int kRowsUnroll = 4;
template <bool SCATTER>
__global__ void __launch_bounds__(256) tree_rows_kernel(uint32_t n_slots, uint32_t f, const int32_t* __restrict__ parents,
                                                        const int32_t* __restrict__ children, int64_t n_graph_nodes,
                                                        int2* __restrict__ segmap, int32_t* __restrict__ rows,
                                                        int32_t* __restrict__ ctr, uint32_t* __restrict__ srcs, int32_t* err) {
    __shared__ int s_cnt, s_base;
    __shared__ int32_t s_stage[256 * kRowsUnroll];  // first-touched rows of one block iteration (<= one per lane and unroll step)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t kPerBlock = 256u * kRowsUnroll;
    const uint32_t n_iter = (n_slots + kPerBlock - 1) / kPerBlock;
    for (uint32_t it = blockIdx.x; it < n_iter; it += gridDim.x) {  // block-uniform
        if (!SCATTER) {
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
        }
        const uint32_t wbase = it * kPerBlock + (uint32_t)warp * (32u * kRowsUnroll);
        int32_t c[kRowsUnroll], v[kRowsUnroll];
        uint32_t p[kRowsUnroll];
#pragma unroll
        for (int u = 0; u < kRowsUnroll; ++u) {
            const uint32_t s = wbase + u * 32u + lane;
            const bool in = s < n_slots;
            p[u] = (in ? s : 0u) / f;
            c[u] = in ? __ldg(children + s) : -1;
        }
#pragma unroll
        for (int u = 0; u < kRowsUnroll; ++u) v[u] = __ldg(parents + p[u]);
        int base[kRowsUnroll];
        uint32_t vm[kRowsUnroll];
        int l0[kRowsUnroll];
        bool first[kRowsUnroll];
#pragma unroll
        for (int u = 0; u < kRowsUnroll; ++u) {
            const uint32_t w0 = wbase + u * 32u;
            const bool in = w0 + lane < n_slots;
            bool valid = in && c[u] >= 0 && v[u] >= 0;
            if (valid && ((int64_t)c[u] >= n_graph_nodes || (int64_t)v[u] >= n_graph_nodes)) {
                atomicExch(err, GIGL_E_RANGE);
                valid = false;
            }
            // the lanes of my parent inside this group of 32 slots: [l0, l1)
            const int64_t seg_lo = (int64_t)p[u] * f - (int64_t)w0;
            l0[u] = seg_lo < 0 ? 0 : (int)seg_lo;
            const int l1 = seg_lo + f > 32 ? 32 : (int)(seg_lo + f);
            uint32_t segmask = (l1 - l0[u] >= 32) ? 0xffffffffu : (((1u << (l1 > l0[u] ? l1 - l0[u] : 0)) - 1u) << l0[u]);
            if (!in) segmask = 0;
            vm[u] = __ballot_sync(0xffffffffu, valid) & segmask;  // the filled slots of my parent inside this group
            const int cnt = __popc(vm[u]);
            base[u] = 0;
            first[u] = false;
            if (in && lane == l0[u] && cnt > 0) {
                if (!SCATTER)
                    first[u] = atomicAdd(&segmap[v[u]].x, cnt) == 0;
                else
                    base[u] = atomicAdd(&segmap[v[u]].y, cnt);
            }
        }
        if (!SCATTER) {
            // rows touched for the first time: staged per block, ONE global cursor atomic per block iteration (a cursor
            // atomic per warp made 3e5 same-address atomics the critical path of the launch)
#pragma unroll
            for (int u = 0; u < kRowsUnroll; ++u) {
                const uint32_t fm = __ballot_sync(0xffffffffu, first[u]);
                if (fm) {
                    int pos = 0;
                    if (lane == __ffs(fm) - 1) pos = atomicAdd(&s_cnt, __popc(fm));
                    pos = __shfl_sync(0xffffffffu, pos, __ffs(fm) - 1);
                    if (first[u]) s_stage[pos + __popc(fm & ((1u << lane) - 1u))] = v[u];
                }
            }
            __syncthreads();
            const int n_staged = s_cnt;
            if (threadIdx.x == 0 && n_staged > 0) s_base = atomicAdd(ctr + kCtrRows, n_staged);
            __syncthreads();
            for (int i = threadIdx.x; i < n_staged; i += 256) rows[s_base + i] = s_stage[i];
            __syncthreads();
        } else {
#pragma unroll
            for (int u = 0; u < kRowsUnroll; ++u) {
                const int b = __shfl_sync(0xffffffffu, base[u], l0[u]);
                if (vm[u] >> lane & 1u) srcs[b + __popc(vm[u] & ((1u << lane) - 1u))] = (uint32_t)c[u];
            }
        }
    }
}

__global__ void __launch_bounds__(256) rows_alloc_kernel(const int32_t* __restrict__ rows, int2* __restrict__ segmap,
                                                         int32_t* __restrict__ ctr, int4* __restrict__ long_items, int long_cap) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_rows = ctr[kCtrRows];
    for (int i0 = blockIdx.x * 256; i0 < n_rows; i0 += gridDim.x * 256) {  // block-uniform
        const int i = i0 + threadIdx.x;
        int v = -1, c = 0;
        if (i < n_rows) {
            v = rows[i];
            c = segmap[v].x;
        }
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int t = s_warp[w];
                s_warp[w] = tot;
                tot += t;
            }
            s_base = tot ? atomicAdd(ctr + kCtrValid, tot) : 0;
        }
        __syncthreads();
        if (i < n_rows) {
            const int beg = s_base + s_warp[warp] + incl - c;
            segmap[v] = make_int2(beg, beg);  // .y = fill cursor of the scatter, ends as the row's end
            if (c > kWarpSortMax) {
                const int R = c > kLongCap ? (c + kLongPart - 1) / kLongPart : 1;
                const int slot = atomicAdd(ctr + kCtrLong, R);
                for (int k = 0; k < R; ++k)
                    if (slot + k < long_cap) long_items[slot + k] = make_int4(v, k, R, c);
            }
        }
        __syncthreads();
    }
}

// Bitonic sort of R * 32 values held by a warp, element r * 32 + lane in e[r] of lane `lane`, ascending.
template <int R>
__device__ __forceinline__ void bitonic_warp(uint32_t (&e)[R], int lane) {
#pragma unroll
    for (int k = 2; k <= R * 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int rj = j >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rj) == 0) {
                        const bool up = ((r * 32) & k) == 0;
                        const uint32_t a = e[r], b = e[r | rj];
                        const uint32_t lo = min(a, b), hi = max(a, b);
                        e[r] = up ? lo : hi;
                        e[r | rj] = up ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, e[r], j);
                    const bool up = (((r * 32) | lane) & k) == 0;
                    const bool take_min = up == ((lane & j) == 0);
                    e[r] = take_min ? min(e[r], o) : max(e[r], o);
                }
            }
        }
    }
}

// Loads a row of c <= R * 32 sources, sorts it, writes its keys; returns this lane's count of first occurrences.
template <int R>
__device__ __forceinline__ int sort_row_warp(const uint32_t* __restrict__ src, uint64_t* __restrict__ dst, int c, uint64_t hi_bits,
                                             int lane) {
    uint32_t e[R];
#pragma unroll
    for (int r = 0; r < R; ++r) e[r] = (r * 32 + lane < c) ? src[r * 32 + lane] : 0xFFFFFFFFu;
    bitonic_warp<R>(e, lane);
    int uniq = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint32_t prev = __shfl_up_sync(0xffffffffu, e[r], 1);
        if (r > 0) {
            const uint32_t carry = __shfl_sync(0xffffffffu, e[r - 1], 31);
            if (lane == 0) prev = carry;
        }
        const int i = r * 32 + lane;
        if (i < c) {
            dst[i] = hi_bits | e[r];
            uniq += (i == 0) || (prev != e[r]);
        }
    }
    return uniq;
}

// The per-row sorts of a batch, one launch.  Every CTA first takes LONG-row work items (row v, partition k of R) from a
// queue: it reads the whole row, keeps the sources inside its value range [lo, hi] in shared memory and counts those
// below lo - that count IS the output offset of its sorted run, so the partitions of one row are independent work items.
// A range holding more than kLongCap sources is split in four and redone (a width-1 range is a run of one repeated
// source); R = 1 for rows that fit as a whole.  Then its warps take TILES of 32 listed rows: lane l resolves row l's
// vertex and bounds (one round trip for 32 rows), the first 32 sources of every row are loaded into 32 registers per
// lane (32 independent loads in flight), and ONE bitonic network - 15 stages, each a shuffle and a min/max per register -
// sorts the 32 rows side by side: 15 compare-exchanges per row of <= 32 entries, with no dependent stall between
// them.  Rows of 33 .. kWarpSortMax entries of the tile are then sorted one at a time in 2 .. 16 registers per lane.
__device__ __forceinline__ void sort_long_item(const int4 it, const int2* __restrict__ segmap, const uint32_t* __restrict__ srcs,
                                               uint64_t* __restrict__ keys, int shift, uint32_t* s_buf, uint2* s_stack, int* s_sp,
                                               int* s_cnt, int* s_below, int& uniq) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int2 seg = segmap[it.x];
    const int c = seg.y - seg.x;
    const uint32_t* src = srcs + seg.x;
    uint64_t* dst = keys + seg.x;
    const uint64_t hi_bits = (uint64_t)(uint32_t)it.x << shift;
    const uint64_t space = 1ULL << shift;
    if (tid == 0) {
        // ranges are [lo, hi] inclusive (hi = 2^32 - 1 must be representable)
        s_stack[0] = make_uint2((uint32_t)(space * (uint64_t)it.y / (uint64_t)it.z), (uint32_t)(space * (uint64_t)(it.y + 1) / (uint64_t)it.z - 1));
        *s_sp = 1;
    }
    __syncthreads();
    while (*s_sp > 0) {  // block-uniform: s_sp only changes between barriers
        const uint2 rg = s_stack[*s_sp - 1];
        __syncthreads();
        if (tid == 0) {
            *s_sp -= 1;
            *s_cnt = 0;
            *s_below = 0;
        }
        __syncthreads();
        int below = 0;
        for (int e0 = 0; e0 < c; e0 += kLongThreads) {
            const int e = e0 + tid;
            const uint32_t x = e < c ? src[e] : 0xFFFFFFFFu;
            const bool inr = e < c && x >= rg.x && x <= rg.y;
            below += e < c && x < rg.x;
            const uint32_t m = __ballot_sync(0xffffffffu, inr);
            if (m) {
                int pos = 0;
                if (lane == __ffs(m) - 1) pos = atomicAdd(s_cnt, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
                if (inr && pos < kLongCap) s_buf[pos] = x;
            }
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) below += __shfl_xor_sync(0xffffffffu, below, off);
        if (lane == 0 && below) atomicAdd(s_below, below);
        __syncthreads();
        const int n = *s_cnt, off0 = *s_below;
        if (n <= kLongCap) {
            int m = 32;
            while (m < n) m <<= 1;
            for (int i = n + tid; i < m; i += kLongThreads) s_buf[i] = 0xFFFFFFFFu;
            __syncthreads();
            for (int k = 2; k <= m; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = tid; t < (m >> 1); t += kLongThreads) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const uint32_t a = s_buf[i], b = s_buf[i | j];
                        if ((a > b) == ((i & k) == 0)) {
                            s_buf[i] = b;
                            s_buf[i | j] = a;
                        }
                    }
                    __syncthreads();
                }
            for (int i = tid; i < n; i += kLongThreads) {
                const uint32_t x = s_buf[i];
                dst[off0 + i] = hi_bits | x;
                uniq += (i == 0) || (s_buf[i - 1] != x);
            }
        } else if (rg.x == rg.y) {
            for (int i = tid; i < n; i += kLongThreads) dst[off0 + i] = hi_bits | rg.x;
            uniq += tid == 0;
        } else if (tid == 0) {
            const uint64_t w = (uint64_t)rg.y - rg.x + 1;
            const int parts = w < 4 ? (int)w : 4;
            for (int q = 0; q < parts; ++q)
                s_stack[(*s_sp)++] = make_uint2(rg.x + (uint32_t)(w * q / parts), rg.x + (uint32_t)(w * (q + 1) / parts - 1));
        }
        __syncthreads();
    }
}

constexpr int kCtrLongNext = 17, kCtrTileNext = 18;

constexpr int kSubTile = 8;  // rows sorted side by side per pass of a tile (registers per lane)

__global__ void __launch_bounds__(kLongThreads, 2) rows_sort_kernel(const int32_t* __restrict__ rows, const int2* __restrict__ segmap,
                                                                    int32_t* __restrict__ ctr, const int4* __restrict__ long_items,
                                                                    int long_cap, const uint32_t* __restrict__ srcs,
                                                                    uint64_t* __restrict__ keys, int shift) {
    __shared__ uint32_t s_buf[kLongCap];  // a long item's range
    __shared__ uint2 s_stack[64];
    __shared__ int s_sp, s_cnt, s_below, s_item;
    const int tid = threadIdx.x, lane = tid & 31;
    int uniq = 0;
    // ---- long rows first (the longest single work items of the launch)
    int n_items = ctr[kCtrLong];
    if (n_items > long_cap) n_items = long_cap;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(ctr + kCtrLongNext, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= n_items) break;
        sort_long_item(long_items[item], segmap, srcs, keys, shift, s_buf, s_stack, &s_sp, &s_cnt, &s_below, uniq);
    }
    // ---- tiles of 32 listed rows per warp
    const int n_rows = ctr[kCtrRows];
    const int n_tiles = (n_rows + 31) >> 5;
    for (;;) {
        int tile = 0;
        if (lane == 0) tile = atomicAdd(ctr + kCtrTileNext, 1);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= n_tiles) break;
        const int i = tile * 32 + lane;
        int32_t v = 0;
        int2 seg = make_int2(0, 0);
        if (i < n_rows) {
            v = __ldg(rows + i);
            seg = segmap[v];
        }
        const int c = seg.y - seg.x;
        // rows of <= 32 sources: kSubTile rows side by side, one register each; the 15 stages of one bitonic network sort
        // them all (a shuffle and a min / max per register and stage, the kSubTile chains independent)
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += kSubTile) {
            uint32_t e[kSubTile];
#pragma unroll
            for (int r = 0; r < kSubTile; ++r) {
                const int cr = __shfl_sync(0xffffffffu, c, r0 + r), br = __shfl_sync(0xffffffffu, seg.x, r0 + r);
                e[r] = (cr <= 32 && lane < cr) ? srcs[br + lane] : 0xFFFFFFFFu;
            }
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
#pragma unroll
                    for (int r = 0; r < kSubTile; ++r) {
                        const uint32_t o = __shfl_xor_sync(0xffffffffu, e[r], j);
                        e[r] = take_min ? min(e[r], o) : max(e[r], o);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kSubTile; ++r) {
                const int cr = __shfl_sync(0xffffffffu, c, r0 + r), br = __shfl_sync(0xffffffffu, seg.x, r0 + r);
                const int32_t vr = __shfl_sync(0xffffffffu, v, r0 + r);
                const uint32_t prev = __shfl_up_sync(0xffffffffu, e[r], 1);
                if (cr <= 32 && lane < cr) {
                    keys[br + lane] = ((uint64_t)(uint32_t)vr << shift) | e[r];
                    uniq += (lane == 0) || (prev != e[r]);
                }
            }
        }
        // the tile's rows of 33 .. kWarpSortMax sources, one at a time in 2 .. 16 registers per lane
        uint32_t mid = __ballot_sync(0xffffffffu, c > 32 && c <= kWarpSortMax);
        while (mid) {
            const int r = __ffs(mid) - 1;
            mid &= mid - 1;
            const int cr = __shfl_sync(0xffffffffu, c, r), br = __shfl_sync(0xffffffffu, seg.x, r);
            const uint64_t hi_bits = (uint64_t)(uint32_t)__shfl_sync(0xffffffffu, v, r) << shift;
            if (cr <= 64)
                uniq += sort_row_warp<2>(srcs + br, keys + br, cr, hi_bits, lane);
            else if (cr <= 128)
                uniq += sort_row_warp<4>(srcs + br, keys + br, cr, hi_bits, lane);
            else if (cr <= 256)
                uniq += sort_row_warp<8>(srcs + br, keys + br, cr, hi_bits, lane);
            else
                uniq += sort_row_warp<16>(srcs + br, keys + br, cr, hi_bits, lane);
        }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) uniq += __shfl_xor_sync(0xffffffffu, uniq, off);
    if (lane == 0 && uniq) atomicAdd(ctr + kCtrUnique, uniq);
}

__global__ void rows_clear_kernel(const int32_t* __restrict__ ctr, const int32_t* __restrict__ rows, int2* __restrict__ segmap) {
    const int n = ctr[kCtrRows];
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) segmap[rows[i]] = make_int2(0, 0);
}

// ---- 3. local ids, level by level ------------------------------------------------------------
// level_end[0] = 0, level_end[1] = n_roots, level_end[j + 1] = nodes after the j-th expansion
__global__ void roots_assign_kernel(int64_t n_roots, const int32_t* __restrict__ roots, int64_t n_nodes,
                                    int32_t* __restrict__ lid, int32_t* __restrict__ list, int32_t* err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_roots) return;
    const int32_t v = roots[i];
    if (v < 0 || v >= n_nodes) {
        atomicExch(err, GIGL_E_RANGE);
        list[i] = 0;
        return;
    }
    list[i] = v;
    atomicMin(lid + v, (int32_t)i);  // duplicate roots: the first slot owns the local id
}

// New local ids are claimed with atomicCAS on the dense map, but handed out in bulk: winners are staged in shared
// memory and the block takes ONE range from the global counter per 1024 keys (a single hot counter with one atomic per
// node serialises in the L2 atomic unit).  The work is cut by KEY, not by row: a thread takes sorted keys and asks
// whether the key's destination sits on the level being expanded (lid[dst] in [lo, hi); neighbouring keys share the
// destination, so the lookup is a cached read).  A row-per-warp walk had the longest row of the batch - a hub that is
// the hop-1 neighbour of thousands of roots - as its critical path (170 us measured; 25 us of key traffic).
constexpr int kStageCap = 1024;
constexpr int kExpandKpt = kStageCap / 256;  // keys per thread per block iteration
__global__ void __launch_bounds__(256) expand_level_kernel(const int32_t* __restrict__ level_end, int level, int64_t n_keys,
                                                           const int32_t* __restrict__ n_keys_dev,
                                                           const uint64_t* __restrict__ keys, int shift, uint64_t src_mask,
                                                           int32_t* __restrict__ lid, int32_t* __restrict__ list,
                                                           int32_t* __restrict__ n_nodes_ctr) {
    __shared__ int s_cnt, s_base;
    __shared__ int32_t s_stage[kStageCap];
    const int lane = threadIdx.x & 31;
    if (n_keys_dev != nullptr && *n_keys_dev < n_keys) n_keys = *n_keys_dev;  // the host passed an upper bound
    const int32_t lo = level_end[level - 1], hi = level_end[level];
    const int64_t step = (int64_t)gridDim.x * kStageCap;
    for (int64_t base = (int64_t)blockIdx.x * kStageCap; base < n_keys; base += step) {  // block-uniform trip count
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kExpandKpt; ++u) {
            const int64_t e = base + u * 256 + threadIdx.x;
            int32_t sv = -1;
            bool won = false;
            if (e < n_keys) {
                const uint64_t k = keys[e];
                const int32_t ld = lid[(uint32_t)(k >> shift)];  // ids of this level were all assigned by earlier launches
                if (ld >= lo && ld < hi) {
                    sv = (int32_t)(k & src_mask);
                    if (lid[sv] == kLidAbsent) won = atomicCAS(lid + sv, kLidAbsent, kLidPending) == kLidAbsent;
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, won);
            if (m == 0) continue;
            int slot = 0;
            if (lane == 0) slot = atomicAdd(&s_cnt, __popc(m));
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (won) s_stage[slot + __popc(m & ((1u << lane) - 1u))] = sv;  // <= kStageCap winners per iteration by construction
        }
        __syncthreads();
        const int n_staged = s_cnt;
        if (threadIdx.x == 0 && n_staged > 0) s_base = atomicAdd(n_nodes_ctr, n_staged);
        __syncthreads();
        for (int i = threadIdx.x; i < n_staged; i += blockDim.x) {
            const int32_t sv = s_stage[i];
            list[s_base + i] = sv;
            lid[sv] = s_base + i;
        }
        __syncthreads();
    }
}

__global__ void level_snapshot_kernel(int32_t* level_end, int level, const int32_t* n_nodes_ctr) {
    level_end[level] = *n_nodes_ctr;
}

__global__ void lid_clear_kernel(const int32_t* __restrict__ n_nodes_ctr, const int32_t* __restrict__ list,
                                 int32_t* __restrict__ lid) {
    const int64_t n = *n_nodes_ctr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) lid[list[i]] = kLidAbsent;
}

__global__ void fill_i32_kernel(int64_t n, int32_t v, int32_t* __restrict__ p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// ---- 4. gather: [mean of unique in-neighbours | self] rows -----------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

constexpr int kSplitThreshold = 256;  // rows longer than this are split into parts ...
constexpr int kPartEdges = 64;        // ... of this many sorted keys, one warp each (32 for rows wider than 128 floats:
                                      // a part is a chain of dependent load batches, and wide rows take two column passes)
constexpr int kGatherUnroll = 8;      // independent 16-byte loads in flight per lane

// Sum of the unique sources of keys[e_beg, e_end) (a slice of the row that starts at row_beg) for
// feature columns [c, c+4) of this lane; every lane group g walks its own subset, the caller
// reduces across groups.  n_uniq counts the slice's unique keys (warp-uniform).
template <int LPR>
__device__ __forceinline__ void gather_range(int row_beg, int e_beg, int e_end, int c, bool active,
                                             const uint64_t* __restrict__ keys, uint64_t src_mask,
                                             const float* __restrict__ xsrc, int64_t ldx,
                                             const int32_t* __restrict__ lid, int lane, int g, float4& acc,
                                             int& n_uniq) {
    constexpr int G = 32 / LPR;
    for (int base = e_beg; base < e_end; base += 32) {
        const int cnt = min(32, e_end - base);
        int32_t my = -1;
        if (lane < cnt) {
            const uint64_t k = __ldg(keys + base + lane);
            const bool dup = (base + lane > row_beg) && (__ldg(keys + base + lane - 1) == k);
            if (!dup) {
                const int32_t s = (int32_t)(k & src_mask);
                my = lid ? __ldg(lid + s) : s;
            }
        }
        n_uniq += __popc(__ballot_sync(0xffffffffu, my >= 0));
        for (int t = 0; t < cnt; t += kGatherUnroll * G) {
            float4 val[kGatherUnroll];
#pragma unroll
            for (int u = 0; u < kGatherUnroll; ++u) {
                const int j = t + u * G + g;
                const int32_t s = __shfl_sync(0xffffffffu, my, j & 31);
                val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < cnt && s >= 0 && active) val[u] = ldg4(xsrc + (int64_t)s * ldx + c);
            }
#pragma unroll
            for (int u = 0; u < kGatherUnroll; ++u) {
                acc.x += val[u].x;
                acc.y += val[u].y;
                acc.z += val[u].z;
                acc.w += val[u].w;
            }
        }
    }
}

// The projection runs as 3xTF32 on the tensor cores (gemm_tcgen05.cu): its A operand is stored
// already split, hi = the value rounded to TF32, lo = (v - hi) rounded to TF32 (gigl_split_tf32).
__device__ __forceinline__ void store_split4(float* __restrict__ hi, float* __restrict__ lo, int64_t off, float4 v) {
    if (lo == nullptr) {  // raw operand: the projection splits it in shared memory (gemm_tcgen05.cu, split_a)
        *reinterpret_cast<float4*>(hi + off) = v;
        return;
    }
    float4 h, l;
    gigl_split_tf32(v.x, h.x, l.x);
    gigl_split_tf32(v.y, h.y, l.y);
    gigl_split_tf32(v.z, h.z, l.z);
    gigl_split_tf32(v.w, h.w, l.w);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

template <int LPR>
__device__ __forceinline__ void reduce_groups(float4& acc) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
    }
}

// Accumulate form of the gather's output (project-first layers, see batch_sage_forward): instead of writing
// [mean | self] rows for a projection, the mean of the gathered rows is ADDED to out[row, 0 .. n_cols) (+ ReLU).
struct AccOut {
    float* out;  // nullptr = the [mean | self] form
    int64_t ldo;
    int32_t n_cols;
    int32_t relu;
};

__device__ __forceinline__ void acc_row4(const AccOut& ao, int64_t row, int c, float4 m) {
    const float v[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (c + j < ao.n_cols) {
            float* o = ao.out + row * ao.ldo + c + j;
            const float t = *o + v[j];
            *o = ao.relu ? fmaxf(t, 0.f) : t;
        }
    }
}

// Work lists of the split path: hctr[0] = parts claimed, hctr[1] = heavy rows claimed.
struct HeavyLists {
    int32_t* hctr;
    int2* items;       // [part_cap]  (row, part index)
    int4* rows;        // [part_cap]  (row, first item, n_parts, -)
    float* partial;    // [part_cap, F]
    int32_t* pcnt;     // [part_cap]
    int32_t part_cap;
    int32_t part_edges;  // sorted keys per part
};

// Sum of the already-resolved sources of one 32-key chunk (`my` per lane, -1 = skip).
template <int LPR>
__device__ __forceinline__ void gather_chunk(int32_t my, int cnt, int c, bool active, const float* __restrict__ xsrc,
                                             int64_t ldx, int g, float4& acc) {
    constexpr int G = 32 / LPR;
    for (int t = 0; t < cnt; t += kGatherUnroll * G) {
        float4 val[kGatherUnroll];
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            const int j = t + u * G + g;
            const int32_t s = __shfl_sync(0xffffffffu, my, j & 31);
            val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < cnt && s >= 0 && active) val[u] = ldg4(xsrc + (int64_t)s * ldx + c);
        }
#pragma unroll
        for (int u = 0; u < kGatherUnroll; ++u) {
            acc.x += val[u].x;
            acc.y += val[u].y;
            acc.z += val[u].z;
            acc.w += val[u].w;
        }
    }
}

// LPR lanes cover one 4*LPR-float chunk of a feature row; G = 32/LPR sources are in flight per step.
// xsrc rows are indexed by global vertex id when lid == nullptr (layer 1 reads the graph-wide
// feature table directly) or by lid[src] (deeper layers read the previous layer's rows).
//
// Persistent warps, one row per iteration.  A row needs a chain of dependent random loads before
// its first feature byte moves (node id -> row bounds -> sorted keys -> local id), ~2 us of pure
// latency; the chain is software-pipelined across iterations: while row i is gathered, the node
// id of row i+4, the bounds of row i+3, the keys of row i+2 and the local ids of row i+1 are in flight.
template <int LPR>
__global__ void __launch_bounds__(256) batch_gather_kernel(const int32_t* __restrict__ n_rows_dev, int64_t row_cap,
                                                           int F, const int32_t* __restrict__ list,
                                                           const int2* __restrict__ segmap,
                                                           const uint64_t* __restrict__ keys, uint64_t src_mask,
                                                           const float* __restrict__ xsrc, int64_t ldx,
                                                           const int32_t* __restrict__ lid, float* __restrict__ A_hi,
                                                           float* __restrict__ A_lo, int64_t ldA, const HeavyLists hl,
                                                           const AccOut ao) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, g = lane / LPR;
    int64_t n_rows = *n_rows_dev;
    if (n_rows > row_cap) n_rows = row_cap;
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t last = n_rows - 1;
#define GIGL_ROW(r) ((r) < n_rows ? (r) : last)
    // first chunk of a row: key per lane (kKeyNone beyond the row), duplicates of the lower lane masked later
    constexpr uint64_t kKeyNone = ~0ULL;
    auto load_keys = [&](const int2& seg) -> uint64_t {
        const int e = seg.x + lane;
        return e < seg.y ? __ldg(keys + e) : kKeyNone;
    };
    auto resolve = [&](uint64_t k) -> int32_t {  // -1 for absent / duplicate, else the source row index
        const uint64_t kp = __shfl_up_sync(0xffffffffu, k, 1);
        if (k == kKeyNone || (lane > 0 && kp == k)) return -1;
        const int32_t sv = (int32_t)(k & src_mask);
        return lid ? __ldg(lid + sv) : sv;
    };
    int32_t vA = __ldg(list + row), vB = __ldg(list + GIGL_ROW(row + W)), vC = __ldg(list + GIGL_ROW(row + 2 * W)),
            vD = __ldg(list + GIGL_ROW(row + 3 * W));
    int2 segA = __ldg(segmap + vA), segB = __ldg(segmap + vB), segC = __ldg(segmap + vC);
    uint64_t kB = load_keys(segB);
    int32_t myA = resolve(load_keys(segA));
    // the row's own source row: through the same map as its neighbours' (the staged table of the halo is not in local-id order)
    int32_t sA = lid ? __ldg(lid + vA) : vA;
    for (; row < n_rows; row += W) {
        // ---- prefetch stage of the pipeline ----
        const int32_t sB = lid ? __ldg(lid + vB) : vB;
        const int32_t vE = __ldg(list + GIGL_ROW(row + 4 * W));
        const int2 segD = __ldg(segmap + vD);
        const uint64_t kC = load_keys(segC);
        const int32_t myB = resolve(kB);
        // ---- row A ----
        const int len = segA.y - segA.x;
        bool deferred = false;
        if (len > kSplitThreshold && hl.items != nullptr) {
            const int n_parts = (len + hl.part_edges - 1) / hl.part_edges;
            int slot = 0;
            if (lane == 0) slot = atomicAdd(hl.hctr, n_parts);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot + n_parts <= hl.part_cap) {
                if (lane == 0) hl.rows[atomicAdd(hl.hctr + 1, 1)] = make_int4((int)row, slot, n_parts, 0);
                for (int p = lane; p < n_parts; p += 32) hl.items[slot + p] = make_int2((int)row, p);
                deferred = true;
            }
            // else: lists full (cannot happen with the host-side sizing): do the long row here
        }
        if (!deferred) {
            const int64_t self = sA;
            const int cnt0 = len < 32 ? len : 32;
            for (int c0 = 0; c0 < F; c0 += LPR * 4) {
                const int c = c0 + sub * 4;
                const bool active = c < F;
                float4 selfv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g == 0 && active && ao.out == nullptr) selfv = ldg4(xsrc + self * ldx + c);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                int n_uniq = __popc(__ballot_sync(0xffffffffu, myA >= 0));
                gather_chunk<LPR>(myA, cnt0, c, active, xsrc, ldx, g, acc);
                if (len > 32)
                    gather_range<LPR>(segA.x, segA.x + 32, segA.y, c, active, keys, src_mask, xsrc, ldx, lid, lane, g, acc, n_uniq);
                reduce_groups<LPR>(acc);
                if (g == 0 && active) {
                    const float scale = 1.0f / (float)(n_uniq > 1 ? n_uniq : 1);
                    const float4 mean = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
                    if (ao.out != nullptr) {
                        acc_row4(ao, row, c, mean);
                    } else {
                        store_split4(A_hi, A_lo, row * ldA + c, mean);
                        store_split4(A_hi, A_lo, row * ldA + F + c, selfv);
                    }
                }
            }
        }
        // ---- rotate ----
        vA = vB;
        vB = vC;
        vC = vD;
        vD = vE;
        segA = segB;
        segB = segC;
        segC = segD;
        kB = kC;
        myA = myB;
        sA = sB;
    }
#undef GIGL_ROW
}

// ---- cp.async-staged gather (feature rows wider than 64 floats) ------------------------------
// Same row walk and metadata pipeline as batch_gather_kernel, but the neighbour rows are not staged
// in registers: every lane copies ITS 16-byte column slice of each neighbour row straight into a
// private shared-memory ring with cp.async (LDGSTS, L2-only), two buffers of SB rows per warp, and
// sums it back from there.  A lane only ever reads bytes it copied itself, so no warp barrier is
// needed, the bytes in flight are bounded by shared memory (8 KB per warp) instead of registers,
// and the first unit of the NEXT row is already in flight while the last unit of a row is summed.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int CPL, int SB>
__global__ void __launch_bounds__(256) batch_gather_async_kernel(const int32_t* __restrict__ n_rows_dev, int64_t row_cap,
                                                                 int F, const int32_t* __restrict__ list,
                                                                 const int2* __restrict__ segmap,
                                                                 const uint64_t* __restrict__ keys, uint64_t src_mask,
                                                                 const float* __restrict__ xsrc, int64_t ldx,
                                                                 const int32_t* __restrict__ lid,
                                                                 float* __restrict__ A_hi, float* __restrict__ A_lo,
                                                                 int64_t ldA, const HeavyLists hl) {
    extern __shared__ float4 s_ring_all[];
    constexpr int kSlot = CPL * 32;  // float4 per staged row
    const int lane = threadIdx.x & 31;
    float4* ring = s_ring_all + (size_t)(threadIdx.x >> 5) * (2 * SB * kSlot);
    int64_t n_rows = *n_rows_dev;
    if (n_rows > row_cap) n_rows = row_cap;
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t last = n_rows - 1;
#define GIGL_ROW(r) ((r) < n_rows ? (r) : last)
    constexpr uint64_t kKeyNone = ~0ULL;
    auto load_keys = [&](int beg, int end) -> uint64_t {
        const int e = beg + lane;
        return e < end ? __ldg(keys + e) : kKeyNone;
    };
    // -1 for absent / duplicate, else the source row index; `prev` = the key just below this chunk
    auto resolve = [&](uint64_t k, uint64_t prev) -> int32_t {
        uint64_t kp = __shfl_up_sync(0xffffffffu, k, 1);
        if (lane == 0) kp = prev;
        if (k == kKeyNone || kp == k) return -1;
        const int32_t sv = (int32_t)(k & src_mask);
        return lid ? __ldg(lid + sv) : sv;
    };
    // copies unit [e0, e0 + SB) of a chunk whose per-lane sources are `my` (chunk starts at entry c_beg)
    auto issue = [&](int buf, int32_t my, int e0, int len) {
#pragma unroll
        for (int jj = 0; jj < SB; ++jj) {
            const int e = e0 + jj;
            const int32_t sv = __shfl_sync(0xffffffffu, my, e & 31);
            if (e < len) {
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int c = (q * 32 + lane) * 4;
                    if (c < F)
                        cp_async16(ring + (buf * SB + jj) * kSlot + q * 32 + lane, xsrc + (sv >= 0 ? (int64_t)sv * ldx : 0) + c,
                                   sv >= 0 ? 16 : 0);
                }
            }
        }
        cp_async_commit();
    };
    int32_t vA = __ldg(list + row), vB = __ldg(list + GIGL_ROW(row + W)), vC = __ldg(list + GIGL_ROW(row + 2 * W)),
            vD = __ldg(list + GIGL_ROW(row + 3 * W));
    int2 segA = __ldg(segmap + vA), segB = __ldg(segmap + vB), segC = __ldg(segmap + vC);
    uint64_t kB = load_keys(segB.x, segB.y);
    int32_t myA = resolve(load_keys(segA.x, segA.y), kKeyNone);
    int32_t sA = lid ? __ldg(lid + vA) : vA;  // the row's own source row, through the same map as its neighbours'
    int ub = 0;              // buffer the next unit of the current row goes to / comes from
    bool pre_issued = false; // unit 0 of row A is already in flight in buffer `ub`
    for (; row < n_rows; row += W) {
        // ---- prefetch stage of the metadata pipeline ----
        const int32_t sB = lid ? __ldg(lid + vB) : vB;
        const int32_t vE = __ldg(list + GIGL_ROW(row + 4 * W));
        const int2 segD = __ldg(segmap + vD);
        const uint64_t kC = load_keys(segC.x, segC.y);
        const int32_t myB = resolve(kB, kKeyNone);
        const int lenB = (row + W < n_rows) ? segB.y - segB.x : 0;
        const bool b_inline = lenB > 0 && !(lenB > kSplitThreshold && hl.items != nullptr);
        // ---- row A ----
        const int len = segA.y - segA.x;
        bool deferred = false;
        if (len > kSplitThreshold && hl.items != nullptr) {
            const int n_parts = (len + hl.part_edges - 1) / hl.part_edges;
            int slot = 0;
            if (lane == 0) slot = atomicAdd(hl.hctr, n_parts);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot + n_parts <= hl.part_cap) {
                if (lane == 0) hl.rows[atomicAdd(hl.hctr + 1, 1)] = make_int4((int)row, slot, n_parts, 0);
                for (int p = lane; p < n_parts; p += 32) hl.items[slot + p] = make_int2((int)row, p);
                deferred = true;
            }
        }
        bool next_pre = false;
        if (!deferred) {
            const int64_t self = sA;
            float4 selfv[CPL], acc[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int c = (q * 32 + lane) * 4;
                selfv[q] = c < F ? ldg4(xsrc + self * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            int n_uniq = 0;
            const int n_units = (len + SB - 1) / SB;
            int32_t my = myA;
            if (n_units > 0) n_uniq = __popc(__ballot_sync(0xffffffffu, my >= 0));
            if (n_units > 0 && !pre_issued) issue(ub, my, 0, len);
            for (int u = 0; u < n_units; ++u) {
                // put the next unit in flight (of this row, or unit 0 of the next row) before summing this one
                bool more = false;
                int32_t my_next = my;
                if (u + 1 < n_units) {
                    const int e0 = (u + 1) * SB;
                    if ((e0 & 31) == 0) {  // next unit starts a new 32-key chunk of a long row
                        const uint64_t prev = __ldg(keys + segA.x + e0 - 1);
                        my_next = resolve(load_keys(segA.x + e0, segA.y), prev);
                        n_uniq += __popc(__ballot_sync(0xffffffffu, my_next >= 0));
                    }
                    issue(ub ^ 1, my_next, e0, len);
                    more = true;
                } else if (b_inline) {
                    issue(ub ^ 1, myB, 0, lenB);
                    more = true;
                    next_pre = true;
                }
                if (more)
                    cp_async_wait<1>();
                else
                    cp_async_wait<0>();
                const int cnt = min(SB, len - u * SB);
#pragma unroll
                for (int jj = 0; jj < SB; ++jj) {
                    if (jj < cnt) {
#pragma unroll
                        for (int q = 0; q < CPL; ++q) {
                            if ((q * 32 + lane) * 4 < F) {
                                const float4 t = ring[(ub * SB + jj) * kSlot + q * 32 + lane];
                                acc[q].x += t.x;
                                acc[q].y += t.y;
                                acc[q].z += t.z;
                                acc[q].w += t.w;
                            }
                        }
                    }
                }
                my = my_next;
                ub ^= 1;
            }
            const float scale = 1.0f / (float)(n_uniq > 1 ? n_uniq : 1);
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int c = (q * 32 + lane) * 4;
                if (c < F) {
                    store_split4(A_hi, A_lo, row * ldA + c, make_float4(acc[q].x * scale, acc[q].y * scale, acc[q].z * scale, acc[q].w * scale));
                    store_split4(A_hi, A_lo, row * ldA + F + c, selfv[q]);
                }
            }
        }
        pre_issued = next_pre;
        // ---- rotate ----
        vA = vB;
        vB = vC;
        vC = vD;
        vD = vE;
        segA = segB;
        segB = segC;
        segC = segD;
        kB = kC;
        myA = myB;
        sA = sB;
    }
#undef GIGL_ROW
}

// One warp per (row, part): partial sums of kPartEdges sorted keys.
template <int LPR>
__global__ void __launch_bounds__(256) batch_gather_parts_kernel(int F, const int32_t* __restrict__ list,
                                                                 const int2* __restrict__ segmap,
                                                                 const uint64_t* __restrict__ keys, uint64_t src_mask,
                                                                 const float* __restrict__ xsrc, int64_t ldx,
                                                                 const int32_t* __restrict__ lid, const HeavyLists hl) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, g = lane / LPR;
    int n_items = hl.hctr[0];
    if (n_items > hl.part_cap) n_items = hl.part_cap;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < n_items; it += warps) {
        const int2 item = hl.items[it];
        const int2 seg = __ldg(segmap + __ldg(list + item.x));
        const int e_beg = seg.x + item.y * hl.part_edges;
        const int e_end = min(seg.y, e_beg + hl.part_edges);
        for (int c0 = 0; c0 < F; c0 += LPR * 4) {
            const int c = c0 + sub * 4;
            const bool active = c < F;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int n_uniq = 0;
            gather_range<LPR>(seg.x, e_beg, e_end, c, active, keys, src_mask, xsrc, ldx, lid, lane, g, acc, n_uniq);
            reduce_groups<LPR>(acc);
            if (g == 0 && active) *reinterpret_cast<float4*>(hl.partial + (int64_t)it * F + c) = acc;
            if (lane == 0 && c0 == 0) hl.pcnt[it] = n_uniq;
        }
    }
}

// One CTA per split row: warp w sums parts w, w + 8, ... in order, the eight warp sums are added in warp order
// (a fixed tree, so the result does not depend on scheduling), then [mean | self].  A hub row has thousands of parts:
// one warp walking them alone was a 40 us critical path for a few KB of work.
constexpr int kFinishWarps = 8;
__global__ void __launch_bounds__(kFinishWarps * 32) batch_gather_finish_kernel(int F, const float* __restrict__ xsrc, int64_t ldx,
                                                                              const int32_t* __restrict__ list,
                                                                              const int32_t* __restrict__ lid,
                                                                              float* __restrict__ A_hi, float* __restrict__ A_lo,
                                                                              int64_t ldA, const HeavyLists hl, const AccOut ao) {
    __shared__ float4 s_acc[kFinishWarps][32];
    __shared__ int s_cnt[kFinishWarps];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n_heavy = hl.hctr[1];
    for (int hr = blockIdx.x; hr < n_heavy; hr += gridDim.x) {
        const int4 r = hl.rows[hr];
        const int64_t row = r.x;
        const int32_t vrow = __ldg(list + row);
        const int64_t self = lid ? (int64_t)__ldg(lid + vrow) : (int64_t)vrow;
        int cnt = 0;
        for (int p = w * 32 + lane; p < r.z; p += kFinishWarps * 32) cnt += hl.pcnt[r.y + p];
#pragma unroll
        for (int off = 16; off; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) s_cnt[w] = cnt;
        for (int c0 = 0; c0 < F; c0 += 128) {
            const int c = c0 + lane * 4;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < F) {
                for (int p = w; p < r.z; p += kFinishWarps) {
                    const float4 t = *reinterpret_cast<const float4*>(hl.partial + (int64_t)(r.y + p) * F + c);
                    acc.x += t.x;
                    acc.y += t.y;
                    acc.z += t.z;
                    acc.w += t.w;
                }
            }
            s_acc[w][lane] = acc;
            __syncthreads();
            if (w == 0 && c < F) {
                int n_uniq = 0;
#pragma unroll
                for (int k = 0; k < kFinishWarps; ++k) n_uniq += s_cnt[k];
                const float scale = 1.0f / (float)(n_uniq > 1 ? n_uniq : 1);
#pragma unroll
                for (int k = 1; k < kFinishWarps; ++k) {
                    const float4 t = s_acc[k][lane];
                    acc.x += t.x;
                    acc.y += t.y;
                    acc.z += t.z;
                    acc.w += t.w;
                }
                const float4 mean = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
                if (ao.out != nullptr) {
                    acc_row4(ao, row, c, mean);
                } else {
                    store_split4(A_hi, A_lo, row * ldA + c, mean);
                    store_split4(A_hi, A_lo, row * ldA + F + c, ldg4(xsrc + self * ldx + c));
                }
            }
            __syncthreads();
        }
    }
}

// Any F / alignment.
__global__ void __launch_bounds__(256) batch_gather_scalar_kernel(const int32_t* __restrict__ n_rows_dev,
                                                                  int64_t row_cap, int F,
                                                                  const int32_t* __restrict__ list,
                                                                  const int2* __restrict__ segmap,
                                                                  const uint64_t* __restrict__ keys, uint64_t src_mask,
                                                                  const float* __restrict__ xsrc, int64_t ldx,
                                                                  const int32_t* __restrict__ lid,
                                                                  float* __restrict__ A_hi, float* __restrict__ A_lo,
                                                                  int64_t ldA, const AccOut ao) {
    const int lane = threadIdx.x & 31;
    int64_t n_rows = *n_rows_dev;
    if (n_rows > row_cap) n_rows = row_cap;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows; row += warps) {
        const int32_t v = __ldg(list + row);
        const int2 seg = __ldg(segmap + v);
        const int64_t self = lid ? (int64_t)__ldg(lid + v) : (int64_t)v;
        for (int c = lane; c < F; c += 32) {
            float acc = 0.f;
            int n_uniq = 0;
            for (int e = seg.x; e < seg.y; ++e) {
                const uint64_t k = __ldg(keys + e);
                if (e > seg.x && __ldg(keys + e - 1) == k) continue;
                const int32_t s0 = (int32_t)(k & src_mask);
                const int64_t s = lid ? __ldg(lid + s0) : s0;
                acc += __ldg(xsrc + s * ldx + c);
                ++n_uniq;
            }
            const float m = acc * (1.0f / (float)(n_uniq > 1 ? n_uniq : 1));
            if (ao.out != nullptr) {
                if (c < ao.n_cols) {
                    float* o = ao.out + row * ao.ldo + c;
                    const float t = *o + m;
                    *o = ao.relu ? fmaxf(t, 0.f) : t;
                }
                continue;
            }
            const float sv = __ldg(xsrc + self * ldx + c);
            if (A_lo == nullptr) {
                A_hi[row * ldA + c] = m;
                A_hi[row * ldA + F + c] = sv;
            } else {
                gigl_split_tf32(m, A_hi[row * ldA + c], A_lo[row * ldA + c]);
                gigl_split_tf32(sv, A_hi[row * ldA + F + c], A_lo[row * ldA + F + c]);
            }
        }
    }
}


// ---- remote-neighbour feature halo -------------------------------------------------------------------
// With the feature table sharded over the box (shared_table.cu) a batch node's row may live in a peer's HBM.  The
// layer-1 gather reads one source row per unique EDGE; staging reads one row per unique NODE: xb[l, :] = x[list[l], :]
// for every local id l, local or over NVLink, after which the gather runs on local memory through `lid`.  One warp per
// row, U rows per warp in flight (a remote row is a ~2 us round trip; 148 SMs x 64 warps x 4 rows x 400 B ~ 15 MB
// in flight).
// Hot rows: the rows of the vertices a batch meets most often (highest degree) are replicated on every GPU
// (gigl_batch_set_hot_rows_dev); hot_slot[v] >= 0 names v's row in that local copy, and only the cold tail crosses NVLink.
template <int CPL>
__global__ void __launch_bounds__(256) halo_stage_kernel(const int32_t* __restrict__ n_nodes_dev, int64_t row_cap, int F,
                                                         const int32_t* __restrict__ list, const float* __restrict__ x,
                                                         int64_t ldx, float* __restrict__ xb, int64_t ldb,
                                                         const int32_t* __restrict__ hot_slot, const float* __restrict__ hot,
                                                         int64_t ldh) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    int64_t n = *n_nodes_dev;
    if (n > row_cap) n = row_cap;
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * U; r0 < n; r0 += W * U) {
        int32_t v[U];
        const float* src[U];
        float4 val[U][CPL];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (r0 + u < n) ? __ldg(list + r0 + u) : -1;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            src[u] = x + (int64_t)(v[u] >= 0 ? v[u] : 0) * ldx;
            if (hot_slot != nullptr && v[u] >= 0) {
                const int32_t hs = __ldg(hot_slot + v[u]);
                if (hs >= 0) src[u] = hot + (int64_t)hs * ldh;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int c = (q * 32 + lane) * 4;
                val[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v[u] >= 0 && c < F) val[u][q] = ldg4(src[u] + c);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int c = (q * 32 + lane) * 4;
                if (v[u] >= 0 && c < F) *reinterpret_cast<float4*>(xb + (r0 + u) * ldb + c) = val[u][q];
            }
    }
}

__global__ void __launch_bounds__(256) halo_stage_scalar_kernel(const int32_t* __restrict__ n_nodes_dev, int64_t row_cap, int F,
                                                                const int32_t* __restrict__ list, const float* __restrict__ x,
                                                                int64_t ldx, float* __restrict__ xb, int64_t ldb,
                                                                const int32_t* __restrict__ hot_slot, const float* __restrict__ hot,
                                                                int64_t ldh) {
    int64_t n = *n_nodes_dev;
    if (n > row_cap) n = row_cap;
    const int lane = threadIdx.x & 31;
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += W) {
        const int32_t v = __ldg(list + r);
        const float* src = x + (int64_t)v * ldx;
        if (hot_slot != nullptr) {
            const int32_t hs = __ldg(hot_slot + v);
            if (hs >= 0) src = hot + (int64_t)hs * ldh;
        }
        for (int c = lane; c < F; c += 32) xb[r * ldb + c] = __ldg(src + c);
    }
}

// ---- the same copy on the bulk-copy engine (TMA, 1-D) ------------------------------------------------------------------
// halo_stage_kernel keeps its bytes in flight in REGISTERS, so a copy that saturates NVLink needs every warp slot of
// every SM (measured 1 / 2 / 8 CTAs per SM -> 0.73 / 0.40 / 0.33 ms) and nothing else runs beside it.  Here every lane owns
// one shared-memory slot and one mbarrier and drives rows through them with cp.async.bulk: global (local HBM or a peer's,
// over NVLink) -> slot, then slot -> the staged table.  The bytes in flight live in shared memory, the kernel occupies
// four warps and a handful of registers per SM, and the collation kernels of the same batch / the other batch's kernels
// keep the rest of the SM.
constexpr int kHaloBarBytes = 128 * 8;
__global__ void __launch_bounds__(128) halo_stage_tma_kernel(const int32_t* __restrict__ n_nodes_dev, const int32_t* __restrict__ first_dev,
                                                             int64_t row_cap, uint32_t row_bytes,
                                                             uint32_t slot_bytes, const int32_t* __restrict__ list,
                                                             const float* __restrict__ x, int64_t ldx, float* __restrict__ xb,
                                                             int64_t ldb, const int32_t* __restrict__ hot_slot,
                                                             const float* __restrict__ hot, int64_t ldh) {
    extern __shared__ __align__(128) uint8_t s_halo[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_halo);                      // one mbarrier per lane
    const uint32_t slot = smem_u32(s_halo + kHaloBarBytes + (size_t)threadIdx.x * slot_bytes);
    const uint32_t bar = smem_u32(bars + threadIdx.x);
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    int64_t n = *n_nodes_dev;
    if (n > row_cap) n = row_cap;
    const int64_t first = first_dev ? (int64_t)*first_dev : 0;  // rows [first, n): the ones claimed since the last copy
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t phase = 0;
    for (int64_t r = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
        const int32_t v = __ldg(list + r);
        const float* src = x + (int64_t)v * ldx;
        if (hot_slot != nullptr) {
            const int32_t hs = __ldg(hot_slot + v);
            if (hs >= 0) src = hot + (int64_t)hs * ldh;
        }
        // the slot's previous store must have read it out
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_expect_tx(bar, row_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot), "l"(src),
                     "r"(row_bytes), "r"(bar)
                     : "memory");
        mbar_wait(bar, phase);
        phase ^= 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(xb + r * ldb), "r"(slot), "r"(row_bytes) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Early staging (gigl_batch_set_halo_table_dev): the rows a batch needs are known as soon as it is SAMPLED - the roots and
// every filled tree slot - so their copy does not have to wait for the collation.  This kernel claims a stage slot for
// every distinct vertex of one id array (atomicCAS on a dense map, winners handed out in blocks like expand_level_kernel);
// halo_stage_kernel then copies x[slist[s]] -> xs[s] on a second stream while the collation runs on the first, and layer 1
// gathers through `sslot` instead of the local-id map.
constexpr int kClaimPerThread = 4;
// d_sctr: [0] = stage slots claimed so far, [1] = copied so far (collate-time staging), [2 + k] = the claim count after the
// k-th staged level (level-by-level staging: the copy of level k moves slots [snap k-1, snap k))
constexpr int kStageCtrInts = 2 + GIGL_MAX_HOPS + 2;
__global__ void __launch_bounds__(256) stage_claim_kernel(int64_t n, const int32_t* __restrict__ ids, int64_t n_graph_nodes,
                                                          int32_t* __restrict__ sslot, int32_t* __restrict__ slist,
                                                          int32_t* __restrict__ sctr) {
    __shared__ int s_cnt, s_base;
    __shared__ int32_t s_stage[256 * kClaimPerThread];
    const int lane = threadIdx.x & 31;
    const int64_t per_block = 256 * kClaimPerThread;
    const int64_t n_iter = (n + per_block - 1) / per_block;
    for (int64_t it = blockIdx.x; it < n_iter; it += gridDim.x) {  // block-uniform
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kClaimPerThread; ++u) {
            const int64_t i = it * per_block + u * 256 + threadIdx.x;
            int32_t v = -1;
            bool won = false;
            if (i < n) {
                v = __ldg(ids + i);
                if (v >= 0 && v < n_graph_nodes && sslot[v] == kLidAbsent) won = atomicCAS(sslot + v, kLidAbsent, kLidPending) == kLidAbsent;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, won);
            if (m == 0) continue;
            int slot = 0;
            if (lane == __ffs(m) - 1) slot = atomicAdd(&s_cnt, __popc(m));
            slot = __shfl_sync(0xffffffffu, slot, __ffs(m) - 1);
            if (won) s_stage[slot + __popc(m & ((1u << lane) - 1u))] = v;
        }
        __syncthreads();
        const int n_staged = s_cnt;
        if (threadIdx.x == 0 && n_staged > 0) s_base = atomicAdd(sctr, n_staged);
        __syncthreads();
        for (int i = threadIdx.x; i < n_staged; i += 256) {
            const int32_t v = s_stage[i];
            slist[s_base + i] = v;
            sslot[v] = s_base + i;
        }
        __syncthreads();
    }
}

}  // namespace gigl

// =================================================================================================
struct gigl_batch {
    gigl_ctx* ctx = nullptr;
    int64_t n_graph_nodes = 0;
    int2* segmap = nullptr;   // dense [n_graph_nodes]: row [beg, end) in the sorted keys, {0,0} = no in-edge
    int32_t* lid = nullptr;   // dense [n_graph_nodes]: local id in this batch, kLidAbsent = not in it
    int32_t* d_ctr = nullptr; // [0] valid keys, [1] unique edges, [2] node counter, [4 + j] level_end[j]
    int32_t* h_ctr = nullptr; // pinned mirror
    // per-collate state
    uint64_t* keys = nullptr;   // sorted keys of the current batch (inside `buf`)
    void* buf = nullptr;        // keys_a | keys_b | list | cub temp
    size_t buf_bytes = 0;
    int32_t* list = nullptr;
    int64_t n_slots = 0;
    int64_t list_cap = 0;
    int shift = 32;            // key = dst << shift | src
    uint64_t src_mask = 0xffffffffULL;
    unsigned long long* d_cursor = nullptr;  // valid-key cursor of the keys kernel
    unsigned long long* h_cursor = nullptr;  // pinned mirror
    int n_levels = 0;          // levels the forward needs (= n_layers of the collate call)
    int n_levels_done = 0;     // level_end entries [1..n_levels_done] are valid
    int n_hops = 0;
    int64_t n_roots = 0;
    bool dirty = false;
    int64_t level_end_host[GIGL_MAX_HOPS + 2] = {};
    int64_t n_valid_host = 0, n_unique_host = 0;
    bool halo_staging = false;  // layer 1 reads a per-batch copy of the unique nodes' rows (batch_set_halo_staging)
    // early staging of the halo (batch_set_halo_table): the registered feature table, the dense vertex -> stage slot map, the
    // staged vertices, their counter, the side stream the claim + copy run on while the collation runs on the ctx stream
    const float* halo_x = nullptr;
    int64_t halo_ldx = 0;
    int32_t halo_F = 0;
    int32_t* sslot = nullptr;
    int32_t* slist = nullptr;
    int64_t slist_cap = 0;
    int32_t* d_sctr = nullptr;
    float* xs = nullptr;
    int64_t lds = 0;
    cudaStream_t halo_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_claimed = nullptr, ev_staged = nullptr;
    bool early_staged = false;   // this batch's rows are (being) copied into xs
    bool prestaged = false;      // ... level by level from inside the sampling call (batch_stage_begin / _level / _end)
    bool stage_open = false;
    bool stage_dirty = false;    // sslot holds entries of slist[0 .. *d_sctr)
    int stage_level = 0;         // levels staged so far by the open batch_stage_begin / _level sequence
    const int32_t* hot_slot = nullptr;  // halo staging: dense [n_graph_nodes] map vertex -> row of the replicated hot table, -1 = cold
    const float* hot = nullptr;
    int64_t ldh = 0;
    int32_t hot_F = 0;
    int32_t* rows = nullptr;    // bucketed collation: the batch's distinct destination vertices (inside `buf`)
    bool bucketed = false;      // the current batch was collated by the bucketed path
};

static constexpr int kCtrInts = 32;
static constexpr int kLevelBase = 4;

static unsigned grid1d(gigl_ctx* ctx, int64_t work, int block) {
    int64_t g = ceil_div64(work > 0 ? work : 1, block);
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    return (unsigned)(g < cap ? g : cap);
}

int batch_create(gigl_ctx* ctx, int64_t n_graph_nodes, gigl_batch** out) {
    using namespace gigl;
    gigl_batch* b = new (std::nothrow) gigl_batch();
    if (!b) return gigl_fail(ctx, GIGL_E_NOMEM, "out of host memory");
    b->ctx = ctx;
    b->n_graph_nodes = n_graph_nodes;
    const size_t nn = (size_t)(n_graph_nodes > 0 ? n_graph_nodes : 1);
    cudaError_t e = cudaMalloc(&b->segmap, sizeof(int2) * nn);
    if (e == cudaSuccess) e = cudaMalloc(&b->lid, sizeof(int32_t) * nn);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_ctr, sizeof(int32_t) * kCtrInts);
    if (e == cudaSuccess) e = cudaMallocHost(&b->h_ctr, sizeof(int32_t) * kCtrInts);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_cursor, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMallocHost(&b->h_cursor, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(b->segmap, 0, sizeof(int2) * nn, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_ctr, 0, sizeof(int32_t) * kCtrInts, ctx->stream);
    if (e != cudaSuccess) {
        if (b->segmap) cudaFree(b->segmap);
        if (b->lid) cudaFree(b->lid);
        if (b->d_ctr) cudaFree(b->d_ctr);
        if (b->h_ctr) cudaFreeHost(b->h_ctr);
        if (b->d_cursor) cudaFree(b->d_cursor);
        if (b->h_cursor) cudaFreeHost(b->h_cursor);
        delete b;
        return gigl_cuda_fail(ctx, e, "batch_create");
    }
    fill_i32_kernel<<<grid1d(ctx, n_graph_nodes, 256), 256, 0, ctx->stream>>>(n_graph_nodes, kLidAbsent, b->lid);
    ctx->launches++;
    *out = b;
    return GIGL_OK;
}

static int batch_clear(gigl_batch* b) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    if (!b->dirty) return GIGL_OK;
    if (b->bucketed) {
        rows_clear_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(b->d_ctr, b->rows, b->segmap);
        GIGL_LAUNCHED(ctx);
    } else if (b->n_valid_host > 0) {
        seg_clear_kernel<<<grid1d(ctx, b->n_valid_host, 256), 256, 0, ctx->stream>>>(b->n_valid_host, b->keys, b->shift, b->segmap);
        GIGL_LAUNCHED(ctx);
    }
    lid_clear_kernel<<<grid1d(ctx, b->list_cap, 256), 256, 0, ctx->stream>>>(b->d_ctr + 2, b->list, b->lid);
    GIGL_LAUNCHED(ctx);
    b->dirty = false;
    return GIGL_OK;
}

// the stage map of the previous batch: cleared on the ctx stream once that batch's copy has finished
static int stage_clear(gigl_batch* b) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    if (!b->stage_dirty) return GIGL_OK;
    GIGL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, b->ev_staged, 0));
    lid_clear_kernel<<<grid1d(ctx, b->slist_cap, 256), 256, 0, ctx->stream>>>(b->d_sctr, b->slist, b->sslot);
    GIGL_LAUNCHED(ctx);
    b->stage_dirty = false;
    b->early_staged = false;
    return GIGL_OK;
}

void batch_destroy(gigl_batch* b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->segmap) cudaFree(b->segmap);
    if (b->lid) cudaFree(b->lid);
    if (b->d_ctr) cudaFree(b->d_ctr);
    if (b->h_ctr) cudaFreeHost(b->h_ctr);
    if (b->d_cursor) cudaFree(b->d_cursor);
    if (b->h_cursor) cudaFreeHost(b->h_cursor);
    if (b->buf) cudaFree(b->buf);
    if (b->halo_stream) {
        cudaStreamSynchronize(b->halo_stream);
        cudaStreamDestroy(b->halo_stream);
    }
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_claimed) cudaEventDestroy(b->ev_claimed);
    if (b->ev_staged) cudaEventDestroy(b->ev_staged);
    if (b->sslot) cudaFree(b->sslot);
    if (b->slist) cudaFree(b->slist);
    if (b->d_sctr) cudaFree(b->d_sctr);
    delete b;
}

static int halo_copy_launch(gigl_batch* b, cudaStream_t st, const int32_t* n_dev, int64_t row_cap, const int32_t* list, const float* x,
                            int64_t ldx, int F0, float* xb, int64_t ldb, const int32_t* first_dev = nullptr) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    // The copy is bound by NVLink round trips, not by SM work; GIGL_HALO_CTAS CTAs per SM (8 = every warp slot: measured best,
    // 1 / 2 / 8 -> 0.73 / 0.40 / 0.33 ms for 1.07 M rows, half of them remote)
    static const int halo_ctas = getenv("GIGL_HALO_CTAS") ? atoi(getenv("GIGL_HALO_CTAS")) : 8;
    const unsigned sgrid = (unsigned)(ctx->sm_count * (halo_ctas > 0 ? halo_ctas : 8));
    const int32_t* hs = (b->hot_slot && b->hot_F == F0) ? b->hot_slot : nullptr;
    const bool vec = (F0 % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0) &&
                     (!hs || (((reinterpret_cast<uintptr_t>(b->hot) & 15) == 0) && b->ldh % 4 == 0));
    // bulk-copy form: rows of whole 16-byte units, 16-byte aligned at both ends, a slot per lane within 64 KB per CTA
    static const bool use_tma = !(getenv("GIGL_HALO_TMA") && getenv("GIGL_HALO_TMA")[0] == '0');
    static const int tma_ctas = getenv("GIGL_HALO_TMA_CTAS") ? atoi(getenv("GIGL_HALO_TMA_CTAS")) : 1;
    // GIGL_HALO_TMA_THREADS rows in flight per CTA, a shared-memory slot each: the footprint decides what can co-run on the SM
    static const int tma_threads_env = getenv("GIGL_HALO_TMA_THREADS") ? atoi(getenv("GIGL_HALO_TMA_THREADS")) : 128;
    const int tma_threads = tma_threads_env >= 32 && tma_threads_env <= 128 ? (tma_threads_env & ~31) : 128;
    const uint32_t row_bytes = (uint32_t)F0 * 4u, slot_bytes = (row_bytes + 15u) & ~15u;
    if (use_tma && vec && ldb % 4 == 0 && ((reinterpret_cast<uintptr_t>(xb) & 15) == 0) && slot_bytes <= 512) {
        const size_t shm = kHaloBarBytes + (size_t)tma_threads * slot_bytes;
        GIGL_CUDA(ctx, cudaFuncSetAttribute(halo_stage_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kHaloBarBytes + 128 * 512)));
        halo_stage_tma_kernel<<<(unsigned)(ctx->sm_count * (tma_ctas > 0 ? tma_ctas : 1)), tma_threads, shm, st>>>(n_dev, first_dev, row_cap, row_bytes, slot_bytes, list,
                                                                                                  x, ldx, xb, ldb, hs, b->hot, b->ldh);
        GIGL_LAUNCHED(ctx);
        return GIGL_OK;
    }
    GIGL_CHECK(ctx, first_dev == nullptr, "level-by-level staging needs the bulk-copy form of the halo (rows of at most 512 bytes, 16-byte aligned)");
    if (vec && F0 <= 128)
        halo_stage_kernel<1><<<sgrid, 256, 0, st>>>(n_dev, row_cap, F0, list, x, ldx, xb, ldb, hs, b->hot, b->ldh);
    else if (vec && F0 <= 256)
        halo_stage_kernel<2><<<sgrid, 256, 0, st>>>(n_dev, row_cap, F0, list, x, ldx, xb, ldb, hs, b->hot, b->ldh);
    else if (vec && F0 <= 512)
        halo_stage_kernel<4><<<sgrid, 256, 0, st>>>(n_dev, row_cap, F0, list, x, ldx, xb, ldb, hs, b->hot, b->ldh);
    else
        halo_stage_scalar_kernel<<<sgrid, 256, 0, st>>>(n_dev, row_cap, F0, list, x, ldx, xb, ldb, hs, b->hot, b->ldh);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

// Forks the halo copy of this batch onto the side stream: claim a stage slot per distinct vertex of the roots and of every
// tree level, then copy the rows.  The ctx stream goes on with the collation; batch_sage_forward joins on ev_staged.
static int stage_early(gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts, int32_t n_hops,
                       const int32_t* const* nbr_dev, int64_t n_slots) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    int rc;
    const int64_t cap = (n_roots + n_slots < b->n_graph_nodes) ? n_roots + n_slots : b->n_graph_nodes;
    if (b->slist_cap < cap) {
        GIGL_CUDA(ctx, cudaStreamSynchronize(b->halo_stream));
        if (b->slist) GIGL_CUDA(ctx, cudaFree(b->slist));
        b->slist = nullptr;
        b->slist_cap = 0;
        GIGL_CUDA(ctx, cudaMalloc(&b->slist, sizeof(int32_t) * (size_t)cap));
        b->slist_cap = cap;
    }
    const int F0 = b->halo_F;
    const int64_t ldb = (F0 + 3) & ~3;
    void* pX = nullptr;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_SAVE, sizeof(float) * (size_t)cap * ldb, &pX)) != GIGL_OK) return rc;
    b->xs = (float*)pX;
    b->lds = ldb;
    cudaStream_t hs = b->halo_stream;
    GIGL_CUDA(ctx, cudaEventRecord(b->ev_fork, ctx->stream));   // the sampler's output (and the cleared stage map) are ready
    GIGL_CUDA(ctx, cudaStreamWaitEvent(hs, b->ev_fork, 0));
    int th = gigl_timer_begin_on(ctx, GIGL_T_HALO_STAGE, hs);
    GIGL_CUDA(ctx, cudaMemsetAsync(b->d_sctr, 0, 2 * sizeof(int32_t), hs));
    int64_t width = n_roots;
    const int32_t* level = roots_dev;
    for (int h = 0; h <= n_hops; ++h) {
        if (width > 0) {
            const unsigned grid = grid1d(ctx, ceil_div64(width, kClaimPerThread), 256);
            stage_claim_kernel<<<grid, 256, 0, hs>>>(width, level, b->n_graph_nodes, b->sslot, b->slist, b->d_sctr);
            GIGL_LAUNCHED(ctx);
        }
        if (h < n_hops) {
            width *= fanouts[h];
            level = nbr_dev[h];
        }
    }
    GIGL_CUDA(ctx, cudaEventRecord(b->ev_claimed, hs));  // the tree is not read after this point on the side stream
    if ((rc = halo_copy_launch(b, hs, b->d_sctr, cap, b->slist, b->halo_x, b->halo_ldx, F0, b->xs, ldb)) != GIGL_OK) return rc;
    gigl_timer_end_on(ctx, th, hs);
    GIGL_CUDA(ctx, cudaEventRecord(b->ev_staged, hs));
    b->early_staged = true;
    b->stage_dirty = true;
    return GIGL_OK;
}

// Level-by-level staging, driven by the sampling call (gigl_sample_khop_staged_dev): hop h's rows are claimed and copied on
// the side stream as soon as hop h is sampled - the roots' and hop-1 rows travel under the hop-2 sampling kernel, the hop-2
// rows under the collation.  d_sctr[0] = claimed so far, d_sctr[1] = copied so far.
bool batch_stages_early(const gigl_batch* b, const float* x_dev) {
    return b->halo_staging && b->halo_x != nullptr && b->halo_x == x_dev;
}

int batch_stage_begin(gigl_batch* b, int64_t n_roots, const int32_t* fanouts, int32_t n_hops) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    int rc;
    if ((rc = stage_clear(b)) != GIGL_OK) return rc;
    int64_t n_slots = 0, width = n_roots;
    for (int h = 0; h < n_hops; ++h) {
        width *= fanouts[h];
        n_slots += width;
    }
    const int64_t cap = (n_roots + n_slots < b->n_graph_nodes) ? n_roots + n_slots : b->n_graph_nodes;
    if (b->slist_cap < cap) {
        GIGL_CUDA(ctx, cudaStreamSynchronize(b->halo_stream));
        if (b->slist) GIGL_CUDA(ctx, cudaFree(b->slist));
        b->slist = nullptr;
        b->slist_cap = 0;
        GIGL_CUDA(ctx, cudaMalloc(&b->slist, sizeof(int32_t) * (size_t)cap));
        b->slist_cap = cap;
    }
    const int64_t ldb = (b->halo_F + 3) & ~3;
    void* pX = nullptr;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_SAVE, sizeof(float) * (size_t)cap * ldb, &pX)) != GIGL_OK) return rc;
    b->xs = (float*)pX;
    b->lds = ldb;
    GIGL_CUDA(ctx, cudaMemsetAsync(b->d_sctr, 0, kStageCtrInts * sizeof(int32_t), ctx->stream));
    b->stage_open = true;
    b->stage_dirty = true;
    b->stage_level = 0;
    return GIGL_OK;
}

int batch_stage_args(gigl_batch* b, gigl_stage_args* out) {
    GIGL_CHECK(b->ctx, b->stage_open && out, "batch_stage_begin first");
    out->slot = b->sslot;
    out->list = b->slist;
    out->ctr = b->d_sctr;
    return GIGL_OK;
}

// ids: one level of the tree (or the roots), just written on the ctx stream.
//   claimed == false (default): a claim pass over `ids` and the copy of what it claimed, both on the side stream.
//   claimed == true: the kernel that wrote `ids` claimed the stage slots as it went (khop_sample_launch with
//   gigl_stage_args; the roots get a claim pass on the ctx stream); the next hop's kernel goes on claiming while this level
//   is copied, so the copy works on a snapshot of the claim count taken between the two.  Measured slower (the claims
//   stretch the sampling kernel by more than the pass costs beside it: profiles/r2_multi_gpu.md) - GIGL_STAGE_CLAIM=sampler.
int batch_stage_level(gigl_batch* b, const int32_t* ids_dev, int64_t n, bool claimed) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    GIGL_CHECK(ctx, b->stage_open, "batch_stage_begin first");
    GIGL_CHECK(ctx, b->stage_level <= GIGL_MAX_HOPS, "more staged levels than hops");
    if (n <= 0) return GIGL_OK;
    cudaStream_t st = ctx->stream, hs = b->halo_stream;
    const int k = b->stage_level++;
    const int32_t *n_dev = b->d_sctr, *first_dev = b->d_sctr + 1;
    if (claimed) {
        if (k == 0) {  // the roots: nobody claimed them
            stage_claim_kernel<<<grid1d(ctx, ceil_div64(n, kClaimPerThread), 256), 256, 0, st>>>(n, ids_dev, b->n_graph_nodes, b->sslot, b->slist, b->d_sctr);
            GIGL_LAUNCHED(ctx);
        }
        int32_t* snap = b->d_sctr + 2 + k;
        GIGL_CUDA(ctx, cudaMemcpyAsync(snap, b->d_sctr, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        n_dev = snap;
        first_dev = k > 0 ? snap - 1 : nullptr;
    }
    GIGL_CUDA(ctx, cudaEventRecord(b->ev_fork, st));
    GIGL_CUDA(ctx, cudaStreamWaitEvent(hs, b->ev_fork, 0));
    int th = gigl_timer_begin_on(ctx, GIGL_T_HALO_STAGE, hs);
    if (!claimed) {
        stage_claim_kernel<<<grid1d(ctx, ceil_div64(n, kClaimPerThread), 256), 256, 0, hs>>>(n, ids_dev, b->n_graph_nodes, b->sslot, b->slist, b->d_sctr);
        GIGL_LAUNCHED(ctx);
        GIGL_CUDA(ctx, cudaEventRecord(b->ev_claimed, hs));  // `ids` is not read after this point on the side stream
    }
    int rc = halo_copy_launch(b, hs, n_dev, b->slist_cap, b->slist, b->halo_x, b->halo_ldx, b->halo_F, b->xs, b->lds, first_dev);
    if (rc != GIGL_OK) return rc;
    if (!claimed) GIGL_CUDA(ctx, cudaMemcpyAsync(b->d_sctr + 1, b->d_sctr, sizeof(int32_t), cudaMemcpyDeviceToDevice, hs));  // copied = claimed
    gigl_timer_end_on(ctx, th, hs);
    GIGL_CUDA(ctx, cudaEventRecord(b->ev_staged, hs));
    return GIGL_OK;
}

int batch_stage_end(gigl_batch* b) {
    b->stage_open = false;
    b->early_staged = true;
    b->prestaged = true;
    return GIGL_OK;
}

int batch_collate(gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts, int32_t n_hops,
                  const int32_t* const* nbr_dev, int32_t n_layers, int64_t* level_sizes_host, int64_t* n_edges_host) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    GIGL_CHECK(ctx, n_hops >= 1 && n_hops <= GIGL_MAX_HOPS, "n_hops must be in [1, 8]");
    GIGL_CHECK(ctx, n_layers >= 1 && n_layers <= GIGL_MAX_HOPS, "n_layers must be in [1, 8]");
    GIGL_CHECK(ctx, n_roots >= 0 && n_roots <= 0x3fffffffLL, "bad n_roots");
    GIGL_CHECK(ctx, fanouts && nbr_dev && (roots_dev || n_roots == 0), "null pointer");
    int tc = gigl_timer_begin(ctx, GIGL_T_COLLATE_MAPS);
    int rc = batch_clear(b);
    if (rc == GIGL_OK && !b->prestaged) rc = stage_clear(b);  // prestaged: the sampling call already staged THIS batch
    gigl_timer_end(ctx, tc);
    if (rc != GIGL_OK) return rc;
    int64_t n_slots = 0, width = n_roots;
    for (int h = 0; h < n_hops; ++h) {
        GIGL_CHECK(ctx, fanouts[h] >= 1 && fanouts[h] <= GIGL_MAX_FANOUT, "fanout must be in [1, 128]");
        width *= fanouts[h];
        n_slots += width;
        GIGL_CHECK(ctx, n_slots <= 0x7fffffffLL, "batch exceeds 2^31-1 edge slots; split the roots");
    }
    const int64_t list_cap = (n_roots + n_slots < b->n_graph_nodes + n_roots) ? n_roots + n_slots : b->n_graph_nodes + n_roots;
    b->shift = bits_for64(b->n_graph_nodes);
    b->src_mask = (1ULL << b->shift) - 1ULL;
    const int end_bit = 2 * b->shift;
    if (b->prestaged)
        b->prestaged = false;  // consumed: early_staged stays set for the forward
    else if (b->halo_staging && b->halo_x != nullptr && n_roots > 0 && (rc = stage_early(b, roots_dev, n_roots, fanouts, n_hops, nbr_dev, n_slots)) != GIGL_OK)
        return rc;
    const size_t ns = ((size_t)(n_slots > 0 ? n_slots : 1) + 31) & ~(size_t)31;
    static const bool use_sort = [] {  // GIGL_COLLATE=sort: one global radix sort of the keys (A/B measurements)
        const char* e = getenv("GIGL_COLLATE");
        return e && e[0] == 's';
    }();
    const size_t list_bytes = (sizeof(int32_t) * (size_t)(list_cap > 0 ? list_cap : 1) + 255) & ~(size_t)255;
    cudaStream_t st = ctx->stream;
    int th;
    int64_t n_valid = 0;
    if (!use_sort) {
        // ---- bucketed path: keys | sources | rows | list | long-row work items
        const int64_t parents_total = n_slots > 0 ? n_slots : 1;  // >= parent slots with children
        const int64_t rows_cap = parents_total < b->n_graph_nodes ? parents_total : b->n_graph_nodes;
        const size_t rows_bytes = (sizeof(int32_t) * (size_t)(rows_cap > 0 ? rows_cap : 1) + 255) & ~(size_t)255;
        const int64_t long_cap = n_slots / 64 + 64;
        const size_t need = sizeof(uint64_t) * ns + sizeof(uint32_t) * ns + rows_bytes + list_bytes + sizeof(int4) * (size_t)long_cap + 256;
        if (b->buf_bytes < need) {
            if (b->buf) GIGL_CUDA(ctx, cudaFree(b->buf));
            b->buf = nullptr;
            b->buf_bytes = 0;
            GIGL_CUDA(ctx, cudaMalloc(&b->buf, need + need / 8));
            b->buf_bytes = need + need / 8;
        }
        uint64_t* keys = (uint64_t*)b->buf;
        uint32_t* srcs = (uint32_t*)(keys + ns);
        b->rows = (int32_t*)(srcs + ns);
        b->list = (int32_t*)((char*)b->rows + rows_bytes);
        int4* long_items = (int4*)((char*)b->list + list_bytes);
        b->keys = keys;
        b->n_slots = n_slots;
        b->list_cap = list_cap;
        b->n_roots = n_roots;
        b->dirty = true;
        b->bucketed = true;
        GIGL_CUDA(ctx, cudaMemsetAsync(b->d_ctr, 0, sizeof(int32_t) * kCtrInts, st));
        th = gigl_timer_begin(ctx, GIGL_T_COLLATE_KEYS);
        for (int pass = 0; pass < 2; ++pass) {
            width = n_roots;
            for (int h = 0; h < n_hops; ++h) {
                const int32_t* parents = (h == 0) ? roots_dev : nbr_dev[h - 1];
                width *= fanouts[h];
                if (width > 0) {
                    const unsigned grid = grid1d(ctx, ceil_div64(width, kRowsUnroll), 256);
                    if (pass == 0)
                        tree_rows_kernel<false><<<grid, 256, 0, st>>>((uint32_t)width, (uint32_t)fanouts[h], parents, nbr_dev[h], b->n_graph_nodes,
                                                                     b->segmap, b->rows, b->d_ctr, srcs, ctx->d_err);
                    else
                        tree_rows_kernel<true><<<grid, 256, 0, st>>>((uint32_t)width, (uint32_t)fanouts[h], parents, nbr_dev[h], b->n_graph_nodes,
                                                                    b->segmap, b->rows, b->d_ctr, srcs, ctx->d_err);
                    GIGL_LAUNCHED(ctx);
                }
            }
            if (pass == 0 && n_slots > 0) {
                rows_alloc_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(b->rows, b->segmap, b->d_ctr, long_items, (int)long_cap);
                GIGL_LAUNCHED(ctx);
            }
        }
        gigl_timer_end(ctx, th);
        th = gigl_timer_begin(ctx, GIGL_T_COLLATE_SORT);
        if (n_slots > 0) {
            rows_sort_kernel<<<ctx->sm_count * 2, kLongThreads, 0, st>>>(b->rows, b->segmap, b->d_ctr, long_items, (int)long_cap, srcs, keys, b->shift);
            GIGL_LAUNCHED(ctx);
        }
        gigl_timer_end(ctx, th);
        n_valid = n_slots;  // device-side count in d_ctr[kCtrValid]; the launches below are sized by the bound
        th = gigl_timer_begin(ctx, GIGL_T_COLLATE_MAPS);
    } else {
    size_t temp_bytes = 0;
    GIGL_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, temp_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                                  (int64_t)n_slots, 0, end_bit, ctx->stream));
    const size_t need = sizeof(uint64_t) * 2 * ns + list_bytes + temp_bytes + 256;
    if (b->buf_bytes < need) {
        if (b->buf) GIGL_CUDA(ctx, cudaFree(b->buf));
        b->buf = nullptr;
        b->buf_bytes = 0;
        GIGL_CUDA(ctx, cudaMalloc(&b->buf, need + need / 8));
        b->buf_bytes = need + need / 8;
    }
    uint64_t* keys_a = (uint64_t*)b->buf;
    uint64_t* keys_b = keys_a + ns;
    b->list = (int32_t*)(keys_b + ns);
    void* temp = (char*)b->list + list_bytes;
    b->keys = keys_b;
    b->n_slots = n_slots;
    b->list_cap = list_cap;
    b->n_roots = n_roots;
    b->dirty = true;
    b->bucketed = false;

    // 1. keys: filled slots only, densely (one host read of the count: the sort then skips the empty slots)
    width = n_roots;
    th = gigl_timer_begin(ctx, GIGL_T_COLLATE_KEYS);
    GIGL_CUDA(ctx, cudaMemsetAsync(b->d_cursor, 0, sizeof(unsigned long long), st));
    for (int h = 0; h < n_hops; ++h) {
        const int32_t* parents = (h == 0) ? roots_dev : nbr_dev[h - 1];
        width *= fanouts[h];
        if (width > 0) {
            tree_keys_kernel<<<grid1d(ctx, ceil_div64(width, kKeysPerThread), 256), 256, 0, st>>>(width, fanouts[h], parents, nbr_dev[h], b->shift, keys_a, b->d_cursor);
            GIGL_LAUNCHED(ctx);
        }
    }
    gigl_timer_end(ctx, th);
    GIGL_CUDA(ctx, cudaMemcpyAsync(b->h_cursor, b->d_cursor, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    GIGL_CUDA(ctx, cudaMemsetAsync(b->d_ctr, 0, sizeof(int32_t) * kCtrInts, st));
    GIGL_CUDA(ctx, cudaStreamSynchronize(st));
    n_valid = (int64_t)*b->h_cursor;
    b->n_valid_host = n_valid;
    // 2. sort + row bounds
    if (n_valid > 0) {
        th = gigl_timer_begin(ctx, GIGL_T_COLLATE_SORT);
        GIGL_CUDA(ctx, cub::DeviceRadixSort::SortKeys(temp, temp_bytes, keys_a, keys_b, n_valid, 0, end_bit, st));
        ctx->launches++;
        gigl_timer_end(ctx, th);
    }
    th = gigl_timer_begin(ctx, GIGL_T_COLLATE_MAPS);
    if (n_valid > 0) {
        seg_bounds_kernel<<<grid1d(ctx, n_valid, 256), 256, 0, st>>>(n_valid, keys_b, b->shift, b->segmap, b->d_ctr);
        GIGL_LAUNCHED(ctx);
    }
    }
    // 3. levels: level_end[1] = n_roots, then n_layers - 1 expansions
    if (n_roots > 0) {
        roots_assign_kernel<<<(unsigned)ceil_div64(n_roots, 256), 256, 0, st>>>(n_roots, roots_dev, b->n_graph_nodes, b->lid, b->list, ctx->d_err);
        GIGL_LAUNCHED(ctx);
    }
    const int32_t init[3] = {(int32_t)n_roots, 0, (int32_t)n_roots};
    GIGL_CUDA(ctx, cudaMemcpyAsync(b->d_ctr + 2, &init[0], sizeof(int32_t), cudaMemcpyHostToDevice, st));
    GIGL_CUDA(ctx, cudaMemcpyAsync(b->d_ctr + kLevelBase, &init[1], sizeof(int32_t) * 2, cudaMemcpyHostToDevice, st));
    // the staged halo copies the row of EVERY batch node, so it needs local ids for all of them: expand through the last
    // hop here (one host read for all the sizes) instead of in the forward (a second expansion launch + host read)
    const int n_expand = (b->halo_staging && !b->early_staged) ? (n_hops > n_layers ? n_hops : n_layers) + 1 : n_layers;
    for (int j = 1; j < n_expand; ++j) {
        if (n_valid > 0) {
            expand_level_kernel<<<grid1d(ctx, ceil_div64(n_valid, kExpandKpt), 256), 256, 0, st>>>(
                b->d_ctr + kLevelBase, j, n_valid, b->bucketed ? b->d_ctr + kCtrValid : nullptr, b->keys, b->shift, b->src_mask, b->lid, b->list,
                b->d_ctr + 2);
            GIGL_LAUNCHED(ctx);
        }
        level_snapshot_kernel<<<1, 1, 0, st>>>(b->d_ctr + kLevelBase, j + 1, b->d_ctr + 2);
        GIGL_LAUNCHED(ctx);
    }
    gigl_timer_end(ctx, th);
    b->n_levels = n_layers;
    b->n_levels_done = n_expand;
    b->n_hops = n_hops;
    if (b->early_staged) GIGL_CUDA(ctx, cudaStreamWaitEvent(st, b->ev_claimed, 0));  // the tree may be overwritten after this call
    GIGL_CUDA(ctx, cudaMemcpyAsync(b->h_ctr, b->d_ctr, sizeof(int32_t) * kCtrInts, cudaMemcpyDeviceToHost, st));
    GIGL_CUDA(ctx, gigl_host_wait(ctx, st));
    b->n_unique_host = b->h_ctr[1];
    if (b->bucketed) b->n_valid_host = b->h_ctr[kCtrValid];
    for (int j = 0; j <= n_expand; ++j) b->level_end_host[j] = b->h_ctr[kLevelBase + j];
    if (level_sizes_host)
        for (int j = 0; j < n_layers; ++j) level_sizes_host[j] = b->level_end_host[j + 1];
    if (n_edges_host) *n_edges_host = b->n_unique_host;
    return GIGL_OK;
}


// ---- model: PyG GraphSAGE weights resident on the device -------------------------------------
struct gigl_sage_model {
    gigl_ctx* ctx = nullptr;
    int n_layers = 0;
    int dims[GIGL_MAX_HOPS + 1] = {};
    int64_t ldw[GIGL_MAX_HOPS] = {};      // leading dimension of Wcat[l] (>= 2 * dims[l], multiple of 4)
    float* wcat[GIGL_MAX_HOPS] = {};      // [dims[l+1], ldw]: row o = [lin_l.weight[o, :] | lin_r.weight[o, :] | 0]
    float* w_hi[GIGL_MAX_HOPS] = {};      // TF32 split of wcat (3xTF32 projection, gemm_tcgen05.cu)
    float* w_lo[GIGL_MAX_HOPS] = {};
    float* bias[GIGL_MAX_HOPS] = {};      // lin_l.bias or nullptr
    // project-first form of a layer l >= 1 whose output is much narrower than its input (4 * Fo <= Fi): lin_l and lin_r as
    // separate operands [o4 | Fo, ldk], so the layer runs as  Z = h @ Wl^T  over the previous level's rows, out = h[rows] @
    // Wr^T + b, out += mean_j Z[j]  - the gather then moves Fo floats per edge instead of Fi (see batch_sage_forward)
    bool pf[GIGL_MAX_HOPS] = {};
    int o4[GIGL_MAX_HOPS] = {};
    int64_t ldk[GIGL_MAX_HOPS] = {};
    float* pl_hi[GIGL_MAX_HOPS] = {};
    float* pl_lo[GIGL_MAX_HOPS] = {};
    float* pr_hi[GIGL_MAX_HOPS] = {};
    float* pr_lo[GIGL_MAX_HOPS] = {};
    float* blob = nullptr;
};

static bool layer_projects_first(int l, int Fi, int Fo) {
    // Off unless GIGL_PROJECT_FIRST=1.  Measured on the bench step (products-like, 256 -> 47 last layer): the gather
    // drops from 0.26 to 0.15 ms but the two skinny projections (N = 48) cost 0.19 ms against 0.07 - a wash
    // (profiles/r2g_*), the M = 128 x N = 48 tcgen05 tiles are bound by the per-stage hand-off, not by flops.
    static const bool enabled = [] {
        const char* e = getenv("GIGL_PROJECT_FIRST");
        return e && e[0] == '1';
    }();
    return enabled && l >= 1 && 4 * Fo <= Fi && Fi % 4 == 0 && ((Fo + 3) & ~3) <= 64;
}

int sage_model_create(gigl_ctx* ctx, int32_t n_layers, const int32_t* dims, const float* const* Wl, const float* const* bl,
                      const float* const* Wr, int weights_on_device, gigl_sage_model** out) {
    GIGL_CHECK(ctx, n_layers >= 1 && n_layers <= GIGL_MAX_HOPS && dims && Wl && Wr && out, "bad model arguments");
    size_t total = 0;
    for (int l = 0; l < n_layers; ++l) {
        GIGL_CHECK(ctx, dims[l] >= 1 && dims[l + 1] >= 1 && Wl[l] && Wr[l], "bad layer");
        const size_t ld = ((size_t)2 * dims[l] + 3) & ~(size_t)3;
        total += 3 * (size_t)dims[l + 1] * ld + (((size_t)dims[l + 1] + 3) & ~(size_t)3);
        if (layer_projects_first(l, dims[l], dims[l + 1]))
            total += 3 * (size_t)(2 * ((dims[l + 1] + 3) & ~3)) * (size_t)((dims[l] + 3) & ~3);
    }
    gigl_sage_model* m = new (std::nothrow) gigl_sage_model();
    if (!m) return gigl_fail(ctx, GIGL_E_NOMEM, "out of host memory");
    m->ctx = ctx;
    m->n_layers = n_layers;
    cudaError_t e = cudaMalloc(&m->blob, sizeof(float) * total);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->blob, 0, sizeof(float) * total, ctx->stream);
    size_t off = 0;
    const cudaMemcpyKind kind = weights_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    for (int l = 0; l < n_layers && e == cudaSuccess; ++l) {
        const int Fi = dims[l], Fo = dims[l + 1];
        const size_t ld = ((size_t)2 * Fi + 3) & ~(size_t)3;
        m->dims[l] = Fi;
        m->dims[l + 1] = Fo;
        m->ldw[l] = (int64_t)ld;
        m->wcat[l] = m->blob + off;
        m->w_hi[l] = m->blob + off + (size_t)Fo * ld;
        m->w_lo[l] = m->blob + off + 2 * (size_t)Fo * ld;
        off += 3 * (size_t)Fo * ld;
        e = cudaMemcpy2DAsync(m->wcat[l], sizeof(float) * ld, Wl[l], sizeof(float) * Fi, sizeof(float) * Fi, Fo, kind, ctx->stream);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(m->wcat[l] + Fi, sizeof(float) * ld, Wr[l], sizeof(float) * Fi, sizeof(float) * Fi, Fo, kind, ctx->stream);
        if (bl && bl[l]) {
            m->bias[l] = m->blob + off;
            if (e == cudaSuccess) e = cudaMemcpyAsync(m->bias[l], bl[l], sizeof(float) * Fo, kind, ctx->stream);
        }
        off += ((size_t)Fo + 3) & ~(size_t)3;
        if (layer_projects_first(l, Fi, Fo)) {
            const int o4 = (Fo + 3) & ~3;
            const size_t ldk = ((size_t)Fi + 3) & ~(size_t)3, blk = (size_t)o4 * ldk;
            float* raw_l = m->blob + off;            // [o4, ldk] lin_l.weight, rows Fo .. o4 zero
            float* raw_r = raw_l + blk;              // [o4, ldk] lin_r.weight
            m->pl_hi[l] = raw_r + blk;
            m->pl_lo[l] = m->pl_hi[l] + blk;
            m->pr_hi[l] = m->pl_lo[l] + blk;
            m->pr_lo[l] = m->pr_hi[l] + blk;
            off += 6 * blk;
            m->pf[l] = true;
            m->o4[l] = o4;
            m->ldk[l] = (int64_t)ldk;
            if (e == cudaSuccess) e = cudaMemcpy2DAsync(raw_l, sizeof(float) * ldk, Wl[l], sizeof(float) * Fi, sizeof(float) * Fi, Fo, kind, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpy2DAsync(raw_r, sizeof(float) * ldk, Wr[l], sizeof(float) * Fi, sizeof(float) * Fi, Fo, kind, ctx->stream);
            if (e == cudaSuccess && (split_tf32_launch(ctx, o4, (int)ldk, raw_l, (int64_t)ldk, m->pl_hi[l], m->pl_lo[l], (int64_t)ldk) != GIGL_OK ||
                                     split_tf32_launch(ctx, o4, (int)ldk, raw_r, (int64_t)ldk, m->pr_hi[l], m->pr_lo[l], (int64_t)ldk) != GIGL_OK))
                e = cudaErrorUnknown;
        }
    }
    for (int l = 0; l < n_layers && e == cudaSuccess; ++l)
        if (split_tf32_launch(ctx, dims[l + 1], (int)m->ldw[l], m->wcat[l], m->ldw[l], m->w_hi[l], m->w_lo[l], m->ldw[l]) != GIGL_OK)
            e = cudaErrorUnknown;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // host weight buffers may be freed by the caller
    if (e != cudaSuccess) {
        if (m->blob) cudaFree(m->blob);
        delete m;
        return gigl_cuda_fail(ctx, e, "sage_model_create");
    }
    *out = m;
    return GIGL_OK;
}

void sage_model_destroy(gigl_sage_model* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    if (m->blob) cudaFree(m->blob);
    delete m;
}

int sage_model_dims(const gigl_sage_model* m, int32_t* n_layers, int32_t* dims) {
    if (n_layers) *n_layers = m->n_layers;
    if (dims)
        for (int l = 0; l <= m->n_layers; ++l) dims[l] = m->dims[l];
    return GIGL_OK;
}

// The gather over the first `rows` local ids of the collated batch: [mean | self] rows of width 2 F into A (a projection
// follows), or - ao.out set - the mean ADDED to ao.out (project-first layers).  xsrc rows are indexed by global vertex id
// (lidmap == nullptr: layer 1 on the graph-wide feature table) or through the local-id map.
static int batch_gather_launch(gigl_batch* b, const int32_t* rows_dev, int64_t rows, int F, const float* xsrc, int64_t ldx,
                               const int32_t* lidmap, float* A_hi, float* A_lo, int64_t lda, const gigl::AccOut ao) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    int rc;
    const int wpb = 8;
    const int64_t gfull = ceil_div64(rows, wpb), gcap = (int64_t)ctx->sm_count * 3;  // 80 registers -> 3 CTAs per SM
    const unsigned grid = (unsigned)(gfull < gcap ? gfull : gcap);  // persistent warps (grid-stride over rows)
    if (grid == 0) return GIGL_OK;
    const bool vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(xsrc) & 15) == 0) && (ldx % 4 == 0);
    if (!vec) {
        batch_gather_scalar_kernel<<<grid, wpb * 32, 0, st>>>(rows_dev, rows, F, b->list, b->segmap, b->keys, b->src_mask, xsrc, ldx, lidmap, A_hi,
                                                             A_lo, lda, ao);
        GIGL_LAUNCHED(ctx);
        return GIGL_OK;
    }
    // split-row work lists (sized from the collate's valid-key count; see kSplitThreshold)
    const int part_edges = F > 128 ? kPartEdges / 2 : kPartEdges;
    const int64_t part_cap = b->n_valid_host / part_edges + b->n_valid_host / kSplitThreshold + 16;
    const size_t o_items = 256;
    const size_t o_rows = o_items + ((sizeof(int2) * (size_t)part_cap + 255) & ~(size_t)255);
    const size_t o_pcnt = o_rows + ((sizeof(int4) * (size_t)part_cap + 255) & ~(size_t)255);
    const size_t o_part = o_pcnt + ((sizeof(int32_t) * (size_t)part_cap + 255) & ~(size_t)255);
    const size_t hbytes = o_part + sizeof(float) * (size_t)part_cap * F;
    void* ph = nullptr;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_WORK, hbytes, &ph)) != GIGL_OK) return rc;
    HeavyLists hl;
    hl.hctr = (int32_t*)ph;
    hl.items = (int2*)((char*)ph + o_items);
    hl.rows = (int4*)((char*)ph + o_rows);
    hl.pcnt = (int32_t*)((char*)ph + o_pcnt);
    hl.partial = (float*)((char*)ph + o_part);
    hl.part_cap = (int32_t)part_cap;
    hl.part_edges = part_edges;
    GIGL_CUDA(ctx, cudaMemsetAsync(hl.hctr, 0, 2 * sizeof(int32_t), st));
    const int hgrid = ctx->sm_count * 4;
#define GIGL_GATHER(LPR)                                                                                                  \
    do {                                                                                                                  \
        batch_gather_kernel<LPR><<<grid, wpb * 32, 0, st>>>(rows_dev, rows, F, b->list, b->segmap, b->keys, b->src_mask,  \
                                                            xsrc, ldx, lidmap, A_hi, A_lo, lda, hl, ao);                  \
        GIGL_LAUNCHED(ctx);                                                                                               \
        batch_gather_parts_kernel<LPR><<<hgrid, 256, 0, st>>>(F, b->list, b->segmap, b->keys, b->src_mask, xsrc, ldx, lidmap, hl); \
        GIGL_LAUNCHED(ctx);                                                                                               \
    } while (0)
#define GIGL_GATHER_ASYNC(CPL, SB)                                                                                        \
    do {                                                                                                                  \
        const size_t shm = (size_t)wpb * 2 * SB * CPL * 32 * sizeof(float4);                                              \
        GIGL_CUDA(ctx, cudaFuncSetAttribute(batch_gather_async_kernel<CPL, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); \
        batch_gather_async_kernel<CPL, SB><<<grid, wpb * 32, shm, st>>>(rows_dev, rows, F, b->list, b->segmap, b->keys,   \
                                                                        b->src_mask, xsrc, ldx, lidmap, A_hi, A_lo, lda, hl); \
        GIGL_LAUNCHED(ctx);                                                                                               \
        batch_gather_parts_kernel<32><<<hgrid, 256, 0, st>>>(F, b->list, b->segmap, b->keys, b->src_mask, xsrc, ldx, lidmap, hl); \
        GIGL_LAUNCHED(ctx);                                                                                               \
    } while (0)
    if (F <= 16)
        GIGL_GATHER(4);
    else if (F <= 32)
        GIGL_GATHER(8);
    else if (F <= 64 || ao.out != nullptr)  // the accumulate form exists in the register-staged kernels only
        GIGL_GATHER(16);
    else if (F <= 128)
        GIGL_GATHER_ASYNC(1, 8);
    else if (F <= 256)
        GIGL_GATHER_ASYNC(2, 4);
    else if (F <= 512)
        GIGL_GATHER_ASYNC(4, 2);
    else if (F <= 1024)  // MAG240M-wide rows (F = 769 stored as 772): one 3 KB row per ring slot
        GIGL_GATHER_ASYNC(8, 1);
    else
        GIGL_GATHER(32);
#undef GIGL_GATHER
#undef GIGL_GATHER_ASYNC
    batch_gather_finish_kernel<<<hgrid, 256, 0, st>>>(F, xsrc, ldx, b->list, lidmap, A_hi, A_lo, lda, hl, ao);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

int batch_sage_forward(gigl_batch* b, const gigl_sage_model* m, const float* x_dev, int64_t ldx0, float* out_dev) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    const int n_layers = m->n_layers;
    GIGL_CHECK(ctx, b->dirty && b->n_levels == n_layers, "collate the batch with n_layers == the model's layer count first");
    if (b->n_roots == 0) return GIGL_OK;
    GIGL_CHECK(ctx, x_dev && out_dev && ldx0 >= m->dims[0], "bad feature table");
    cudaStream_t st = ctx->stream;
    size_t a_elems = 0, h_elems = 0;
    for (int l = 1; l <= n_layers; ++l) {
        const int64_t rows = b->level_end_host[n_layers - l + 1];
        const int64_t lda = m->ldw[l - 1];
        if ((size_t)(rows * lda) > a_elems) a_elems = (size_t)(rows * lda);
        if (l < n_layers && (size_t)(rows * m->dims[l]) > h_elems) h_elems = (size_t)(rows * m->dims[l]);
    }
    h_elems = (h_elems + 3) & ~(size_t)3;
    void *pA = nullptr, *pH = nullptr;
    int rc;
    a_elems = (a_elems + 63) & ~(size_t)63;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_AGG, sizeof(float) * a_elems, &pA)) != GIGL_OK) return rc;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO3, sizeof(float) * (2 * h_elems + 8), &pH)) != GIGL_OK) return rc;
    float* A_hi = (float*)pA;  // [mean | self] rows, fp32: the projection splits them into TF32 halves in shared memory
    float* A_lo = nullptr;
    float* hbuf[2] = {(float*)pH, (float*)pH + h_elems};
    const float* xin = x_dev;
    int64_t ldx = ldx0;
    bool staged = false;
    const int32_t* stage_map = nullptr;  // layer 1 reads a staged copy through this map (local ids, or the early stage slots)
    if (b->halo_staging && b->early_staged && b->halo_x == x_dev && b->halo_F == m->dims[0]) {
        // the copy was forked inside the collation (stage_early): join it here
        int tw = gigl_timer_begin(ctx, GIGL_T_HALO_WAIT);
        GIGL_CUDA(ctx, cudaStreamWaitEvent(st, b->ev_staged, 0));
        gigl_timer_end(ctx, tw);
        xin = b->xs;
        ldx = b->lds;
        stage_map = b->sslot;
        staged = true;
    } else if (b->halo_staging) {
        // every batch node gets a local id (the levels above stop at what the root outputs need), then one row per node
        int64_t n_nodes = 0;
        if ((rc = batch_finalize_nodes(b, &n_nodes, nullptr)) != GIGL_OK) return rc;
        const int F0 = m->dims[0];
        const int64_t ldb = (F0 + 3) & ~3;
        void* pX = nullptr;
        if ((rc = gigl_scratch(ctx, GIGL_SLOT_SAVE, sizeof(float) * (size_t)(n_nodes > 0 ? n_nodes : 1) * ldb, &pX)) != GIGL_OK) return rc;
        float* xb = (float*)pX;
        const int32_t* n_dev = b->d_ctr + kLevelBase + b->n_levels_done;
        int th = gigl_timer_begin(ctx, GIGL_T_HALO_STAGE);
        if ((rc = halo_copy_launch(b, st, n_dev, n_nodes, b->list, x_dev, ldx0, F0, xb, ldb)) != GIGL_OK) return rc;
        gigl_timer_end(ctx, th);
        xin = xb;
        ldx = ldb;
        stage_map = b->lid;
        staged = true;
    }
    for (int l = 1; l <= n_layers; ++l) {
        const int Fi = m->dims[l - 1], Fo = m->dims[l];
        const int64_t rows = b->level_end_host[n_layers - l + 1];
        const int32_t* rows_dev = b->d_ctr + kLevelBase + (n_layers - l + 1);
        const int64_t lda = m->ldw[l - 1];
        const int32_t* lidmap = (l == 1) ? (staged ? stage_map : nullptr) : b->lid;
        float* C = (l == n_layers) ? out_dev : hbuf[l & 1];
        if (l >= 2 && m->pf[l - 1] && (reinterpret_cast<uintptr_t>(xin) & 15) == 0 && ldx % 4 == 0) {
            // ---- project first: Z = h @ Wl^T over every row of the previous level, C = h[rows] @ Wr^T + b, then the
            // gather adds mean_j Z[j] - Fo floats per edge instead of Fi (W mean(h) = mean(W h))
            const int o4 = m->o4[l - 1];
            const int64_t ldk = m->ldk[l - 1];
            const int64_t rows_prev = b->level_end_host[n_layers - l + 2];
            void* pZ = nullptr;
            if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO4, sizeof(float) * (size_t)(rows_prev > 0 ? rows_prev : 1) * o4, &pZ)) != GIGL_OK) return rc;
            float* Z = (float*)pZ;
            int tg = gigl_timer_begin(ctx, GIGL_T_GEMM_DEEP);
            rc = linear_tc_launch_ex(ctx, rows_prev, nullptr, o4, Fi, xin, nullptr, ldx, m->pl_hi[l - 1], m->pl_lo[l - 1], ldk, nullptr, Z, o4, 0);
            if (rc == GIGL_OK)
                rc = linear_tc_launch_ex(ctx, rows, nullptr, Fo, Fi, xin, nullptr, ldx, m->pr_hi[l - 1], m->pr_lo[l - 1], ldk, m->bias[l - 1], C, Fo, 0);
            gigl_timer_end(ctx, tg);
            if (rc != GIGL_OK) return rc;
            const AccOut ao{C, (int64_t)Fo, Fo, l < n_layers ? 1 : 0};
            tg = gigl_timer_begin(ctx, GIGL_T_GATHER_DEEP);
            rc = batch_gather_launch(b, rows_dev, rows, o4, Z, o4, b->lid, nullptr, nullptr, 0, ao);
            gigl_timer_end(ctx, tg);
            if (rc != GIGL_OK) return rc;
            xin = C;
            ldx = Fo;
            continue;
        }
        int tg = gigl_timer_begin(ctx, l == 1 ? GIGL_T_GATHER_L1 : GIGL_T_GATHER_DEEP);
        rc = batch_gather_launch(b, rows_dev, rows, Fi, xin, ldx, lidmap, A_hi, A_lo, lda, AccOut{nullptr, 0, 0, 0});
        gigl_timer_end(ctx, tg);
        if (rc != GIGL_OK) return rc;
        gigl_timed tgemm(ctx, l == 1 ? GIGL_T_GEMM_L1 : GIGL_T_GEMM_DEEP);
        rc = linear_tc_launch_ex(ctx, rows, nullptr, Fo, 2 * Fi, A_hi, A_lo, lda, m->w_hi[l - 1], m->w_lo[l - 1], lda, m->bias[l - 1], C, Fo,
                                 l < n_layers ? 1 : 0);
        if (rc != GIGL_OK) return rc;
        xin = C;
        ldx = Fo;
    }
    return GIGL_OK;
}

gigl_ctx* batch_ctx(gigl_batch* b) { return b->ctx; }

void batch_set_halo_staging(gigl_batch* b, bool enabled) { b->halo_staging = enabled; }

int batch_set_halo_table(gigl_batch* b, const float* x_dev, int32_t F, int64_t ldx) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    GIGL_CHECK(ctx, x_dev == nullptr || (F >= 1 && ldx >= F), "bad feature table shape");
    GIGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (b->halo_stream) GIGL_CUDA(ctx, cudaStreamSynchronize(b->halo_stream));
    b->halo_x = x_dev;
    b->halo_F = F;
    b->halo_ldx = ldx;
    if (x_dev == nullptr) return GIGL_OK;
    if (!b->halo_stream) {
        // highest priority: the copy's few CTAs should get SM slots ahead of the collation's grids they run under
        int prio_lo = 0, prio_hi = 0;
        GIGL_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        static const bool halo_prio = !(getenv("GIGL_HALO_PRIORITY") && getenv("GIGL_HALO_PRIORITY")[0] == '0');
        GIGL_CUDA(ctx, cudaStreamCreateWithPriority(&b->halo_stream, cudaStreamNonBlocking, halo_prio ? prio_hi : prio_lo));
        GIGL_CUDA(ctx, cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
        GIGL_CUDA(ctx, cudaEventCreateWithFlags(&b->ev_claimed, cudaEventDisableTiming));
        GIGL_CUDA(ctx, cudaEventCreateWithFlags(&b->ev_staged, cudaEventDisableTiming));
        const size_t nn = (size_t)(b->n_graph_nodes > 0 ? b->n_graph_nodes : 1);
        GIGL_CUDA(ctx, cudaMalloc(&b->sslot, sizeof(int32_t) * nn));
        GIGL_CUDA(ctx, cudaMalloc(&b->d_sctr, kStageCtrInts * sizeof(int32_t)));
        GIGL_CUDA(ctx, cudaMemsetAsync(b->d_sctr, 0, kStageCtrInts * sizeof(int32_t), ctx->stream));
        fill_i32_kernel<<<grid1d(ctx, b->n_graph_nodes, 256), 256, 0, ctx->stream>>>(b->n_graph_nodes, kLidAbsent, b->sslot);
        GIGL_LAUNCHED(ctx);
        GIGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GIGL_OK;
}

int batch_set_hot_rows(gigl_batch* b, const int32_t* hot_slot_dev, const float* hot_dev, int32_t F, int64_t ld) {
    gigl_ctx* ctx = b->ctx;
    GIGL_CHECK(ctx, (hot_slot_dev == nullptr) == (hot_dev == nullptr), "hot_slot and hot table must be given together");
    GIGL_CHECK(ctx, hot_dev == nullptr || (F >= 1 && ld >= F), "bad hot table shape");
    b->hot_slot = hot_slot_dev;
    b->hot = hot_dev;
    b->hot_F = F;
    b->ldh = ld;
    return GIGL_OK;
}

namespace gigl {
// compacted unique keys (dst << 32 | src, global ids) -> edge_index rows in local ids
__global__ void export_edges_kernel(int64_t n_unique, const uint64_t* __restrict__ uniq, int shift, uint64_t src_mask,
                                    const int32_t* __restrict__ lid, int64_t* __restrict__ edge_index) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_unique; i += stride) {
        const uint64_t k = uniq[i];
        edge_index[i] = lid[(uint32_t)(k & src_mask)];
        edge_index[n_unique + i] = lid[(uint32_t)(k >> shift)];
    }
}
}  // namespace gigl

int batch_finalize_nodes(gigl_batch* b, int64_t* n_nodes, int64_t* n_edges) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    GIGL_CHECK(ctx, b->dirty, "collate a batch first");
    cudaStream_t st = ctx->stream;
    const int want = (b->n_hops > b->n_levels ? b->n_hops : b->n_levels) + 1;  // level_end entries [1..want]
    if (b->n_levels_done < want && b->n_roots > 0) {
        for (int j = b->n_levels_done; j < want; ++j) {
            if (b->n_valid_host > 0) {
                expand_level_kernel<<<grid1d(ctx, ceil_div64(b->n_valid_host, kExpandKpt), 256), 256, 0, st>>>(
                    b->d_ctr + kLevelBase, j, b->n_valid_host, nullptr, b->keys, b->shift, b->src_mask, b->lid, b->list, b->d_ctr + 2);
                GIGL_LAUNCHED(ctx);
            }
            level_snapshot_kernel<<<1, 1, 0, st>>>(b->d_ctr + kLevelBase, j + 1, b->d_ctr + 2);
            GIGL_LAUNCHED(ctx);
        }
        b->n_levels_done = want;
        GIGL_CUDA(ctx, cudaMemcpyAsync(b->h_ctr, b->d_ctr, sizeof(int32_t) * kCtrInts, cudaMemcpyDeviceToHost, st));
        GIGL_CUDA(ctx, gigl_host_wait(ctx, st));
        for (int j = 0; j <= want; ++j) b->level_end_host[j] = b->h_ctr[kLevelBase + j];
    }
    if (n_nodes) *n_nodes = b->n_roots > 0 ? b->level_end_host[b->n_levels_done] : 0;
    if (n_edges) *n_edges = b->n_unique_host;
    return GIGL_OK;
}

int batch_export(gigl_batch* b, int32_t* node_ids_dev, int64_t* edge_index_dev) {
    using namespace gigl;
    gigl_ctx* ctx = b->ctx;
    int64_t n_nodes = 0, n_edges = 0;
    int rc = batch_finalize_nodes(b, &n_nodes, &n_edges);
    if (rc != GIGL_OK) return rc;
    cudaStream_t st = ctx->stream;
    if (n_nodes > 0) {
        GIGL_CHECK(ctx, node_ids_dev != nullptr, "null node_ids");
        GIGL_CUDA(ctx, cudaMemcpyAsync(node_ids_dev, b->list, sizeof(int32_t) * (size_t)n_nodes, cudaMemcpyDeviceToDevice, st));
    }
    if (n_edges > 0) {
        GIGL_CHECK(ctx, edge_index_dev != nullptr, "null edge_index");
        // compact the unique keys (CUB select, plumbing) into scratch, then translate to local ids
        void* pu = nullptr;
        size_t tb = 0;
        GIGL_CUDA(ctx, cub::DeviceSelect::Unique(nullptr, tb, b->keys, (uint64_t*)nullptr, (int32_t*)nullptr, (int64_t)b->n_valid_host, st));
        const size_t ub = (sizeof(uint64_t) * (size_t)b->n_valid_host + 255) & ~(size_t)255;
        if ((rc = gigl_scratch(ctx, GIGL_SLOT_SORT, ub + tb + 256, &pu)) != GIGL_OK) return rc;
        uint64_t* uniq = (uint64_t*)pu;
        GIGL_CUDA(ctx, cub::DeviceSelect::Unique((char*)pu + ub, tb, b->keys, uniq, b->d_ctr + 3, (int64_t)b->n_valid_host, st));
        ctx->launches++;
        export_edges_kernel<<<grid1d(ctx, n_edges, 256), 256, 0, st>>>(n_edges, uniq, b->shift, b->src_mask, b->lid, edge_index_dev);
        GIGL_LAUNCHED(ctx);
    }
    return GIGL_OK;
}

