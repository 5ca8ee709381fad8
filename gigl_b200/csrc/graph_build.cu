// graph_build.cu - edge lists -> CSR resident in HBM.
//
//  * graph_from_edges_build : the reference's edge-load rules
//    (scala/subgraph_sampler/src/main/scala/libs/task/pureSpark/SGSPureSparkV1Task.scala:120-286):
//    ids are int32; undirected graphs are de-duplicated as (least, greatest) and unioned
//    (UNION DISTINCT) with their reverse (enforceBidirectionalization :218-258); directed graphs
//    keep duplicates.  Rows come out sorted ascending = array_sort(collect_list(_src_node)) :334-342.
//  * csr_from_coo_launch : PyG edge_index (row 0 = src, row 1 = dst) of a collated batch
//    (python/gigl/src/common/graph_builder/pyg_graph_builder.py:20-69) -> CSR by dst with every
//    row's edges in input order (stable), so the fp32 accumulation order of the aggregate is the
//    order of a sequential index_add_.
//
// Both are one key-building pass, one CUB radix sort (library plumbing, not the hot op) and one
// boundary pass that writes rowptr from the sorted keys.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "common.cuh"

namespace gigl {

constexpr uint64_t kDeadKey = ~0ULL;

__global__ void edge_keys_kernel(int64_t e, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                 int64_t n_nodes, int directed, int by_source, uint64_t* __restrict__ keys,
                                 int32_t* err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += stride) {
        const int32_t s = src[i], d = dst[i];
        if (s < 0 || d < 0 || s >= n_nodes || d >= n_nodes) {
            atomicExch(err, GIGL_E_RANGE);
            keys[i] = kDeadKey;
            continue;
        }
        uint32_t row, c;
        if (!directed) {
            row = (uint32_t)min(s, d);  // (least, greatest)
            c = (uint32_t)max(s, d);
        } else if (by_source) {
            row = (uint32_t)s;
            c = (uint32_t)d;
        } else {
            row = (uint32_t)d;
            c = (uint32_t)s;
        }
        keys[i] = ((uint64_t)row << 32) | c;
    }
}

// (lo,hi) unique pairs -> both orientations; a self loop survives once (UNION is distinct).
__global__ void mirror_keys_kernel(int64_t m, const uint64_t* __restrict__ uniq, uint64_t* __restrict__ out,
                                   unsigned long long* n_dead) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long dead = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint64_t k = uniq[i];
        if (k == kDeadKey) {  // out-of-range edge already reported
            out[2 * i] = kDeadKey;
            out[2 * i + 1] = kDeadKey;
            dead += 2;
            continue;
        }
        const uint32_t lo = (uint32_t)(k >> 32), hi = (uint32_t)k;
        out[2 * i] = k;
        if (lo == hi) {
            out[2 * i + 1] = kDeadKey;
            dead += 1;
        } else {
            out[2 * i + 1] = ((uint64_t)hi << 32) | lo;
        }
    }
    if (dead) atomicAdd(n_dead, dead);
}

// rowptr from sorted 64-bit (row << 32 | col) keys; also splits off the column.
__global__ void rows_from_keys_kernel(int64_t e, const uint64_t* __restrict__ keys, int64_t n_nodes,
                                      int64_t* __restrict__ rowptr, int32_t* __restrict__ col) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= e; i += stride) {
        const int64_t prev = (i == 0) ? -1 : (int64_t)(keys[i - 1] >> 32);
        const int64_t cur = (i == e) ? n_nodes : (int64_t)(keys[i] >> 32);
        for (int64_t v = prev + 1; v <= cur; ++v) rowptr[v] = i;
        if (i < e) col[i] = (int32_t)(uint32_t)keys[i];
    }
}

__global__ void coo_keys_kernel(int64_t e, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                int64_t n, uint32_t* __restrict__ key, int32_t* __restrict__ val, int32_t* err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += stride) {
        const int64_t s = src[i], d = dst[i];
        if (s < 0 || d < 0 || s >= n || d >= n) {
            atomicExch(err, GIGL_E_RANGE);
            key[i] = (uint32_t)(n > 0 ? n - 1 : 0);  // keep the CSR well formed; the error surfaces at sync
            val[i] = 0;
        } else {
            key[i] = (uint32_t)d;
            val[i] = (int32_t)s;
        }
    }
}

__global__ void rows_from_keys32_kernel(int64_t e, const uint32_t* __restrict__ keys, int64_t n,
                                        int64_t* __restrict__ rowptr) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= e; i += stride) {
        const int64_t prev = (i == 0) ? -1 : (int64_t)keys[i - 1];
        const int64_t cur = (i == e) ? n : (int64_t)keys[i];
        for (int64_t v = prev + 1; v <= cur; ++v) rowptr[v] = i;
    }
}

static inline int bits_for(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

static inline unsigned grid_for(gigl_ctx* ctx, int64_t work, int block) {
    int64_t g = ceil_div64(work > 0 ? work : 1, block);
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    return (unsigned)(g < cap ? g : cap);
}

}  // namespace gigl

int csr_from_coo_launch(gigl_ctx* ctx, int64_t n, int64_t e, const int64_t* src, const int64_t* dst,
                        int64_t* rowptr, int32_t* col) {
    using namespace gigl;
    if (e == 0) {
        GIGL_CUDA(ctx, cudaMemsetAsync(rowptr, 0, sizeof(int64_t) * (size_t)(n + 1), ctx->stream));
        return GIGL_OK;
    }
    GIGL_CHECK(ctx, n > 0, "edges given for an empty node set");
    // sort buffers: key_in | key_out | val_in   (val_out = col)
    const size_t ee = ((size_t)e + 63) & ~(size_t)63;
    void* sbuf = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_SORT, sizeof(uint32_t) * 3 * ee, &sbuf);
    if (rc != GIGL_OK) return rc;
    uint32_t* key_in = (uint32_t*)sbuf;
    uint32_t* key_out = key_in + ee;
    int32_t* val_in = (int32_t*)(key_out + ee);
    size_t temp_bytes = 0;
    const int end_bit = bits_for(n);
    GIGL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, key_in, key_out, val_in, col, e, 0, end_bit, ctx->stream));
    void* temp = nullptr;
    rc = gigl_scratch(ctx, GIGL_SLOT_WORK, temp_bytes, &temp);
    if (rc != GIGL_OK) return rc;
    coo_keys_kernel<<<grid_for(ctx, e, 256), 256, 0, ctx->stream>>>(e, src, dst, n, key_in, val_in, ctx->d_err);
    GIGL_LAUNCHED(ctx);
    GIGL_CUDA(ctx, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, key_in, key_out, val_in, col, e, 0, end_bit, ctx->stream));
    ctx->launches += 1;
    rows_from_keys32_kernel<<<grid_for(ctx, e + 1, 256), 256, 0, ctx->stream>>>(e, key_out, n, rowptr);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

int graph_from_edges_build(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src_dev,
                           const int32_t* dst_dev, int32_t directed, int32_t by_source, int64_t** rowptr_dev,
                           int32_t** col_dev, int64_t* n_edges_out) {
    using namespace gigl;
    *rowptr_dev = nullptr;
    *col_dev = nullptr;
    *n_edges_out = 0;
    cudaStream_t st = ctx->stream;
    int64_t* rowptr = nullptr;
    GIGL_CUDA(ctx, cudaMalloc(&rowptr, sizeof(int64_t) * (size_t)(n_nodes + 1)));
    uint64_t *k0 = nullptr, *k1 = nullptr;
    void* temp = nullptr;
    int32_t* col = nullptr;
    unsigned long long* d_cnt = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        if (k0) cudaFree(k0);
        if (k1) cudaFree(k1);
        if (temp) cudaFree(temp);
        if (d_cnt) cudaFree(d_cnt);
    };
    auto fail = [&](cudaError_t e, const char* what) {
        cleanup();
        if (col) cudaFree(col);
        cudaFree(rowptr);
        return gigl_cuda_fail(ctx, e, what);
    };
#define GB_CUDA(call)                                   \
    do {                                                \
        cudaError_t e__ = (call);                       \
        if (e__ != cudaSuccess) return fail(e__, #call); \
    } while (0)

    if (n_edges == 0) {
        GB_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int64_t) * (size_t)(n_nodes + 1), st));
        GB_CUDA(cudaMalloc(&col, sizeof(int32_t)));
        GB_CUDA(cudaStreamSynchronize(st));
        *rowptr_dev = rowptr;
        *col_dev = col;
        return GIGL_OK;
    }
    // the undirected path needs room for 2x the unique pairs
    const size_t cap = (size_t)n_edges * (directed ? 1 : 2);
    GB_CUDA(cudaMalloc(&k0, sizeof(uint64_t) * cap));
    GB_CUDA(cudaMalloc(&k1, sizeof(uint64_t) * cap));
    GB_CUDA(cudaMalloc(&d_cnt, sizeof(unsigned long long) * 2));
    GB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 2, st));
    const int row_bits = bits_for(n_nodes);
    const int end_bit = 32 + row_bits;
    size_t tb_sort = 0, tb_sort2 = 0, tb_uniq = 0;
    GB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb_sort, k0, k1, (int64_t)n_edges, 0, 64, st));
    if (!directed) {
        GB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb_sort2, k0, k1, (int64_t)cap, 0, 64, st));
        GB_CUDA(cub::DeviceSelect::Unique(nullptr, tb_uniq, k1, k0, d_cnt, (int64_t)n_edges, st));
    }
    size_t tb = tb_sort > tb_sort2 ? tb_sort : tb_sort2;
    if (tb_uniq > tb) tb = tb_uniq;
    GB_CUDA(cudaMalloc(&temp, tb > 0 ? tb : 256));

    edge_keys_kernel<<<grid_for(ctx, n_edges, 256), 256, 0, st>>>(n_edges, src_dev, dst_dev, n_nodes, directed, by_source, k0, ctx->d_err);
    ctx->launches++;
    GB_CUDA(cudaGetLastError());
    int64_t e_final = n_edges;
    const uint64_t* sorted = nullptr;
    if (directed) {
        // out-of-range edges became kDeadKey and sort last (all 64 bits are compared)
        GB_CUDA(cub::DeviceRadixSort::SortKeys(temp, tb, k0, k1, (int64_t)n_edges, 0, 64, st));
        ctx->launches++;
        sorted = k1;
    } else {
        GB_CUDA(cub::DeviceRadixSort::SortKeys(temp, tb, k0, k1, (int64_t)n_edges, 0, 64, st));
        GB_CUDA(cub::DeviceSelect::Unique(temp, tb, k1, k0, d_cnt, (int64_t)n_edges, st));
        ctx->launches += 2;
        unsigned long long m = 0;
        GB_CUDA(cudaMemcpyAsync(&m, d_cnt, sizeof(m), cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        mirror_keys_kernel<<<grid_for(ctx, (int64_t)m, 256), 256, 0, st>>>((int64_t)m, k0, k1, d_cnt + 1);
        ctx->launches++;
        GB_CUDA(cudaGetLastError());
        GB_CUDA(cub::DeviceRadixSort::SortKeys(temp, tb, k1, k0, (int64_t)(2 * m), 0, 64, st));
        ctx->launches++;
        unsigned long long dead = 0;
        GB_CUDA(cudaMemcpyAsync(&dead, d_cnt + 1, sizeof(dead), cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        e_final = (int64_t)(2 * m - dead);
        sorted = k0;
    }
    (void)end_bit;
    // deferred range check: dead keys of the directed path are still inside e_final, so test first
    {
        int32_t code = 0;
        GB_CUDA(cudaMemcpyAsync(&code, ctx->d_err, sizeof(code), cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        if (code != 0) {
            cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), st);
            cleanup();
            cudaFree(rowptr);
            return gigl_fail(ctx, GIGL_E_RANGE, "edge endpoint outside [0, n_nodes)");
        }
    }
    GB_CUDA(cudaMalloc(&col, sizeof(int32_t) * (size_t)(e_final > 0 ? e_final : 1)));
    rows_from_keys_kernel<<<grid_for(ctx, e_final + 1, 256), 256, 0, st>>>(e_final, sorted, n_nodes, rowptr, col);
    ctx->launches++;
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaStreamSynchronize(st));
    cleanup();
#undef GB_CUDA
    *rowptr_dev = rowptr;
    *col_dev = col;
    *n_edges_out = e_final;
    return GIGL_OK;
}

// ---- edge-row map: which input edge record hydrates every slot of the CSR ---------------------------------------
// hydrateEdges joins the sampled (src, dst) pairs with the hydrated edge table on (_from, _to)
// (SGSPureSparkV1Task.scala:540-563), so the encoder needs, per CSR slot, the input record that carries its features.
// Same key order as graph_from_edges_build (by destination), with the record index riding along as the sort value:
//   directed   - every record is its own slot; equal (dst, src) keys keep input order (stable radix sort);
//   undirected - one record per (least, greatest) pair survives enforceBidirectionalization (:218-258; the reference's
//                dropDuplicates keeps an arbitrary one, this build keeps the lowest record index) and hydrates both
//                orientations.
namespace gigl {

__global__ void iota32_kernel(int64_t e, int32_t* __restrict__ v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += stride) v[i] = (int32_t)i;
}

__global__ void mirror_pairs_kernel(int64_t m, const uint64_t* __restrict__ uniq, const int32_t* __restrict__ rep,
                                    uint64_t* __restrict__ out_k, int32_t* __restrict__ out_v, unsigned long long* n_dead) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long dead = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const uint64_t k = uniq[i];
        const uint32_t lo = (uint32_t)(k >> 32), hi = (uint32_t)k;
        const bool no_mirror = (k == kDeadKey || lo == hi);  // a self loop survives once (UNION is distinct)
        out_k[2 * i] = k;
        out_v[2 * i] = rep[i];
        out_k[2 * i + 1] = no_mirror ? kDeadKey : (((uint64_t)hi << 32) | lo);
        out_v[2 * i + 1] = rep[i];
        dead += (k == kDeadKey ? 2 : (no_mirror ? 1 : 0));
    }
    if (dead) atomicAdd(n_dead, dead);
}

}  // namespace gigl

int edge_rows_build(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src_dev, const int32_t* dst_dev,
                    int32_t directed, int32_t* rows_dev, int64_t rows_cap, int64_t* n_rows_out) {
    using namespace gigl;
    *n_rows_out = 0;
    if (n_edges == 0) return GIGL_OK;
    cudaStream_t st = ctx->stream;
    const size_t cap = (size_t)n_edges * (directed ? 1 : 2);
    uint64_t *k0 = nullptr, *k1 = nullptr;
    int32_t *v0 = nullptr, *v1 = nullptr;
    void* temp = nullptr;
    unsigned long long* d_cnt = nullptr;
    auto cleanup = [&]() {
        cudaStreamSynchronize(st);
        cudaFree(k0);
        cudaFree(k1);
        cudaFree(v0);
        cudaFree(v1);
        cudaFree(temp);
        cudaFree(d_cnt);
    };
#define ER_CUDA(call)                                \
    do {                                             \
        cudaError_t e__ = (call);                    \
        if (e__ != cudaSuccess) {                    \
            cleanup();                               \
            return gigl_cuda_fail(ctx, e__, #call);  \
        }                                            \
    } while (0)
    ER_CUDA(cudaMalloc(&k0, sizeof(uint64_t) * cap));
    ER_CUDA(cudaMalloc(&k1, sizeof(uint64_t) * cap));
    ER_CUDA(cudaMalloc(&v0, sizeof(int32_t) * cap));
    ER_CUDA(cudaMalloc(&v1, sizeof(int32_t) * cap));
    ER_CUDA(cudaMalloc(&d_cnt, sizeof(unsigned long long) * 2));
    ER_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 2, st));
    size_t tb = 0, tb2 = 0;
    ER_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, (int64_t)cap, 0, 64, st));
    if (!directed) ER_CUDA(cub::DeviceSelect::UniqueByKey(nullptr, tb2, k1, v1, k0, v0, d_cnt, (int64_t)n_edges, st));
    if (tb2 > tb) tb = tb2;
    ER_CUDA(cudaMalloc(&temp, tb > 0 ? tb : 256));
    edge_keys_kernel<<<grid_for(ctx, n_edges, 256), 256, 0, st>>>(n_edges, src_dev, dst_dev, n_nodes, directed, 0, k0, ctx->d_err);
    iota32_kernel<<<grid_for(ctx, n_edges, 256), 256, 0, st>>>(n_edges, v0);
    ctx->launches += 2;
    ER_CUDA(cudaGetLastError());
    ER_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, k0, k1, v0, v1, (int64_t)n_edges, 0, 64, st));
    ctx->launches++;
    int64_t e_final = n_edges;
    const int32_t* rows = v1;
    if (!directed) {
        ER_CUDA(cub::DeviceSelect::UniqueByKey(temp, tb, k1, v1, k0, v0, d_cnt, (int64_t)n_edges, st));
        unsigned long long m = 0;
        ER_CUDA(cudaMemcpyAsync(&m, d_cnt, sizeof(m), cudaMemcpyDeviceToHost, st));
        ER_CUDA(cudaStreamSynchronize(st));
        mirror_pairs_kernel<<<grid_for(ctx, (int64_t)m, 256), 256, 0, st>>>((int64_t)m, k0, v0, k1, v1, d_cnt + 1);
        ER_CUDA(cudaGetLastError());
        ER_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, k1, k0, v1, v0, (int64_t)(2 * m), 0, 64, st));
        ctx->launches += 3;
        unsigned long long dead = 0;  // dead keys (the missing mirror of a self loop) sort last
        ER_CUDA(cudaMemcpyAsync(&dead, d_cnt + 1, sizeof(dead), cudaMemcpyDeviceToHost, st));
        ER_CUDA(cudaStreamSynchronize(st));
        e_final = (int64_t)(2 * m - dead);
        rows = v0;
    }
    int32_t code = 0;
    ER_CUDA(cudaMemcpyAsync(&code, ctx->d_err, sizeof(code), cudaMemcpyDeviceToHost, st));
    ER_CUDA(cudaStreamSynchronize(st));
    if (code != 0) {
        cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), st);
        cleanup();
        return gigl_fail(ctx, GIGL_E_RANGE, "edge endpoint outside [0, n_nodes)");
    }
    if (e_final > rows_cap) {
        cleanup();
        return gigl_fail(ctx, GIGL_E_INVALID, "edge-row buffer smaller than the CSR");
    }
    ER_CUDA(cudaMemcpyAsync(rows_dev, rows, sizeof(int32_t) * (size_t)e_final, cudaMemcpyDeviceToDevice, st));
    ER_CUDA(cudaStreamSynchronize(st));
    cleanup();
#undef ER_CUDA
    *n_rows_out = e_final;
    return GIGL_OK;
}
