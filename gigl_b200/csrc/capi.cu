// capi.cu - the extern "C" surface of libgigl_b200.so (include/gigl_b200.h): context, graph
// residency, COO->CSR conversion and the host-buffer entry points a JNI / ctypes binding calls.
// The kernels live in khop_sample.cu / sage_aggregate.cu / graph_build.cu.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "common.cuh"

// ---- error plumbing -------------------------------------------------------------------------
static thread_local std::string g_null_ctx_err;

int gigl_fail(gigl_ctx* ctx, int code, const std::string& msg) {
    if (ctx)
        ctx->err = msg;
    else
        g_null_ctx_err = msg;
    return code;
}

int gigl_cuda_fail(gigl_ctx* ctx, cudaError_t e, const char* what) {
    std::string m = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return gigl_fail(ctx, e == cudaErrorMemoryAllocation ? GIGL_E_NOMEM : GIGL_E_CUDA, m);
}

int gigl_scratch(gigl_ctx* ctx, int slot, size_t bytes, void** out) {
    if (slot < 0 || slot >= GIGL_SCRATCH_SLOTS) return gigl_fail(ctx, GIGL_E_INVALID, "bad scratch slot");
    if (bytes < 256) bytes = 256;
    if (ctx->scratch_bytes[slot] < bytes) {
        // growth is rare (sizes are sticky); cudaFree synchronises the device, so work already
        // enqueued on the old buffer has finished before it is released
        if (ctx->scratch[slot]) {
            GIGL_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t want = bytes + bytes / 4;
        want = (want + 255) & ~(size_t)255;
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
        if (e != cudaSuccess) {
            want = (bytes + 255) & ~(size_t)255;
            e = cudaMalloc(&ctx->scratch[slot], want);
        }
        if (e != cudaSuccess) return gigl_cuda_fail(ctx, e, "cudaMalloc(scratch)");
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return GIGL_OK;
}

// ---- phase timing ---------------------------------------------------------------------------
static void timer_flush(gigl_ctx* ctx) {
    if (ctx->t_used == 0) return;
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < ctx->t_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->t_pool[i].a, ctx->t_pool[i].b) == cudaSuccess) {
            ctx->t_ms[ctx->t_pool[i].tag] += ms;
            ctx->t_n[ctx->t_pool[i].tag] += 1;
        }
    }
    ctx->t_used = 0;
    cudaGetLastError();
}

int gigl_timer_begin(gigl_ctx* ctx, int tag) {
    if (!ctx->timing) return -1;
    if (ctx->t_used == ctx->t_cap) timer_flush(ctx);
    const int h = ctx->t_used++;
    ctx->t_pool[h].tag = tag;
    cudaEventRecord(ctx->t_pool[h].a, ctx->stream);
    return h;
}

void gigl_timer_end(gigl_ctx* ctx, int handle) {
    if (handle < 0) return;
    cudaEventRecord(ctx->t_pool[handle].b, ctx->stream);
}

int gigl_timer_begin_on(gigl_ctx* ctx, int tag, cudaStream_t stream) {
    if (!ctx->timing) return -1;
    if (ctx->t_used == ctx->t_cap) timer_flush(ctx);
    const int h = ctx->t_used++;
    ctx->t_pool[h].tag = tag;
    cudaEventRecord(ctx->t_pool[h].a, stream);
    return h;
}

void gigl_timer_end_on(gigl_ctx* ctx, int handle, cudaStream_t stream) {
    if (handle < 0) return;
    cudaEventRecord(ctx->t_pool[handle].b, stream);
}

static const char* kTimerNames[GIGL_T_COUNT] = {"sample", "collate_keys", "collate_sort", "collate_maps", "gather_l1",
                                                "gather_deep", "gemm_l1", "gemm_deep", "gather_full", "gemm_full", "halo_stage", "halo_wait"};

static int ctx_check_device_error(gigl_ctx* ctx) {
    GIGL_CUDA(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    GIGL_CUDA(ctx, gigl_host_wait(ctx, ctx->stream));
    const int32_t code = *ctx->h_err;
    if (code != 0) {
        GIGL_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), ctx->stream));
        if (code == GIGL_E_RANGE) return gigl_fail(ctx, GIGL_E_RANGE, "a vertex id outside [0, n_nodes) was met on the device");
        if (code == GIGL_E_OVERFLOW) return gigl_fail(ctx, GIGL_E_OVERFLOW, "a row (times its multiplicity) exceeds 2^31-1 entries");
        return gigl_fail(ctx, code, "device-side error");
    }
    return GIGL_OK;
}

// Shared by the two *_conv_host entry points: stage x / edge_index / weights, COO->CSR, run `layer`.
template <typename LayerFn>
static int conv_host_common(gigl_ctx* ctx, int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* edge_index,
                            const float* x, const float* const* weights, const size_t* weight_elems, int n_weights,
                            float* out, LayerFn layer) {
    GIGL_CHECK(ctx, n >= 0 && n <= 0x7fffffffLL && e >= 0 && F >= 1 && O >= 1, "bad sizes");
    GIGL_CHECK(ctx, n == 0 || (x && out), "null x / out");
    GIGL_CHECK(ctx, e == 0 || edge_index, "null edge_index");
    if (n == 0) return GIGL_OK;
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    void *p_x = nullptr, *p_ei = nullptr, *p_csr = nullptr, *p_w = nullptr, *p_out = nullptr;
    int rc;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO0, sizeof(float) * (size_t)n * F, &p_x)) != GIGL_OK) return rc;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO1, sizeof(int64_t) * 2 * (size_t)(e > 0 ? e : 1), &p_ei)) != GIGL_OK) return rc;
    const size_t rowptr_bytes = (sizeof(int64_t) * (size_t)(n + 1) + 255) & ~(size_t)255;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO2, rowptr_bytes + sizeof(int32_t) * (size_t)(e > 0 ? e : 1), &p_csr)) != GIGL_OK) return rc;
    size_t w_total = 0;
    size_t w_off[4];
    for (int i = 0; i < n_weights; ++i) {
        w_off[i] = w_total;
        w_total += (weight_elems[i] + 3) & ~(size_t)3;
    }
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO3, sizeof(float) * (w_total > 0 ? w_total : 1), &p_w)) != GIGL_OK) return rc;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO4, sizeof(float) * (size_t)n * O, &p_out)) != GIGL_OK) return rc;
    GIGL_CUDA(ctx, cudaMemcpyAsync(p_x, x, sizeof(float) * (size_t)n * F, cudaMemcpyHostToDevice, ctx->stream));
    if (e > 0)
        GIGL_CUDA(ctx, cudaMemcpyAsync(p_ei, edge_index, sizeof(int64_t) * 2 * (size_t)e, cudaMemcpyHostToDevice, ctx->stream));
    const float* w_dev[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int i = 0; i < n_weights; ++i) {
        if (!weights[i]) continue;
        w_dev[i] = (float*)p_w + w_off[i];
        GIGL_CUDA(ctx, cudaMemcpyAsync((void*)w_dev[i], weights[i], sizeof(float) * weight_elems[i], cudaMemcpyHostToDevice, ctx->stream));
    }
    int64_t* rowptr = (int64_t*)p_csr;
    int32_t* col = (int32_t*)((char*)p_csr + rowptr_bytes);
    const int64_t* ei = (const int64_t*)p_ei;
    if ((rc = csr_from_coo_launch(ctx, n, e, ei, ei + e, rowptr, col)) != GIGL_OK) return rc;
    if ((rc = layer(rowptr, col, (const float*)p_x, w_dev, (float*)p_out)) != GIGL_OK) return rc;
    GIGL_CUDA(ctx, cudaMemcpyAsync(out, p_out, sizeof(float) * (size_t)n * O, cudaMemcpyDeviceToHost, ctx->stream));
    return ctx_check_device_error(ctx);
}

extern "C" {

const char* gigl_version(void) { return "gigl_b200 0.2.0 sm_100a"; }

static int ctx_create_common(int device, void* stream, bool own, gigl_ctx** out) {
    if (!out) return gigl_fail(nullptr, GIGL_E_INVALID, "null out pointer");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return gigl_fail(nullptr, GIGL_E_CUDA, "no usable CUDA device (libgigl_b200 has no CPU fallback)");
    if (device < 0 || device >= n_dev) return gigl_fail(nullptr, GIGL_E_INVALID, "device index out of range");
    gigl_ctx* ctx = new (std::nothrow) gigl_ctx();
    if (!ctx) return gigl_fail(nullptr, GIGL_E_NOMEM, "out of host memory");
    ctx->device = device;
    auto bail = [&](cudaError_t err, const char* what) {
        int rc = gigl_cuda_fail(nullptr, err, what);
        delete ctx;
        return rc;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    if (prop.major != 10) {
        delete ctx;
        return gigl_fail(nullptr, GIGL_E_CUDA, "libgigl_b200 is built for sm_100a only; this device is not compute capability 10.x");
    }
    ctx->sm_count = prop.multiProcessorCount;
    if (own) {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
            return bail(e, "cudaStreamCreate");
        ctx->own_stream = true;
    } else {
        ctx->stream = (cudaStream_t)stream;
    }
    if ((e = cudaMalloc(&ctx->d_err, sizeof(int32_t))) != cudaSuccess) return bail(e, "cudaMalloc(err)");
    if ((e = cudaMemset(ctx->d_err, 0, sizeof(int32_t))) != cudaSuccess) return bail(e, "cudaMemset(err)");
    if ((e = cudaMallocHost(&ctx->h_err, sizeof(int32_t))) != cudaSuccess) return bail(e, "cudaMallocHost(err)");
    *ctx->h_err = 0;
    const char* sync_mode = getenv("GIGL_SYNC");
    if (sync_mode && sync_mode[0] == 'b') {
        if ((e = cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess)
            return bail(e, "cudaEventCreate(blocking)");
        ctx->block_sync = true;
    }
    *out = ctx;
    return GIGL_OK;
}

int gigl_ctx_create(int device, gigl_ctx** out) { return ctx_create_common(device, nullptr, true, out); }

int gigl_ctx_create_on_stream(int device, void* cuda_stream, gigl_ctx** out) {
    return ctx_create_common(device, cuda_stream, false, out);
}

void gigl_ctx_destroy(gigl_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
    for (int s = 0; s < GIGL_SCRATCH_SLOTS; ++s)
        if (ctx->scratch[s]) cudaFree(ctx->scratch[s]);
    if (ctx->t_pool) {
        for (int i = 0; i < ctx->t_cap; ++i) {
            cudaEventDestroy(ctx->t_pool[i].a);
            cudaEventDestroy(ctx->t_pool[i].b);
        }
        delete[] ctx->t_pool;
    }
    if (ctx->lad.blob_a) cudaFree(ctx->lad.blob_a);
    if (ctx->lad.blob_b) cudaFree(ctx->lad.blob_b);
    if (ctx->d_err) cudaFree(ctx->d_err);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->copy_ready) cudaEventDestroy(ctx->copy_ready);
    if (ctx->pack_ready) cudaEventDestroy(ctx->pack_ready);
    if (ctx->h_pack_total) cudaFreeHost(ctx->h_pack_total);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int gigl_ctx_set_timing(gigl_ctx* ctx, int32_t enabled) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (enabled && !ctx->t_pool) {
        const int cap = 1024;
        ctx->t_pool = new (std::nothrow) gigl_timer_pair[cap];
        if (!ctx->t_pool) return gigl_fail(ctx, GIGL_E_NOMEM, "out of host memory");
        for (int i = 0; i < cap; ++i) {
            GIGL_CUDA(ctx, cudaEventCreate(&ctx->t_pool[i].a));
            GIGL_CUDA(ctx, cudaEventCreate(&ctx->t_pool[i].b));
            ctx->t_cap = i + 1;
        }
    }
    if (!enabled) timer_flush(ctx);
    ctx->timing = enabled != 0;
    return GIGL_OK;
}

int gigl_ctx_reset_timing(gigl_ctx* ctx) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    timer_flush(ctx);
    for (int i = 0; i < GIGL_T_COUNT; ++i) {
        ctx->t_ms[i] = 0.0;
        ctx->t_n[i] = 0;
    }
    return GIGL_OK;
}

int32_t gigl_timing_num_tags(void) { return GIGL_T_COUNT; }

const char* gigl_timing_tag_name(int32_t tag) { return (tag >= 0 && tag < GIGL_T_COUNT) ? kTimerNames[tag] : ""; }

int gigl_ctx_get_timing(gigl_ctx* ctx, int32_t tag, double* total_ms, int64_t* count) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, tag >= 0 && tag < GIGL_T_COUNT, "bad timing tag");
    timer_flush(ctx);
    if (total_ms) *total_ms = ctx->t_ms[tag];
    if (count) *count = ctx->t_n[tag];
    return GIGL_OK;
}

int gigl_ctx_sync(gigl_ctx* ctx) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return ctx_check_device_error(ctx);
}

const char* gigl_last_error(gigl_ctx* ctx) { return ctx ? ctx->err.c_str() : g_null_ctx_err.c_str(); }

int64_t gigl_ctx_launch_count(gigl_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* gigl_ctx_stream(gigl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---- graph ---------------------------------------------------------------------------------

static int graph_new(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int64_t* rowptr, const int32_t* col,
                     bool owned, gigl_graph** out) {
    gigl_graph* g = new (std::nothrow) gigl_graph();
    if (!g) return gigl_fail(ctx, GIGL_E_NOMEM, "out of host memory");
    g->ctx = ctx;
    g->n_nodes = n_nodes;
    g->n_edges = n_edges;
    g->rowptr = rowptr;
    g->col = col;
    g->owned = owned;
    *out = g;
    return GIGL_OK;
}

int gigl_graph_create_host(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int64_t* rowptr,
                           const int32_t* col, gigl_graph** out) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, out != nullptr, "null out pointer");
    *out = nullptr;
    GIGL_CHECK(ctx, n_nodes >= 0 && n_nodes <= 0x7fffffffLL, "n_nodes must be in [0, 2^31-1]");
    GIGL_CHECK(ctx, n_edges >= 0, "n_edges < 0");
    GIGL_CHECK(ctx, rowptr != nullptr && (col != nullptr || n_edges == 0), "null rowptr / col");
    GIGL_CHECK(ctx, rowptr[0] == 0 && rowptr[n_nodes] == n_edges, "rowptr[0] must be 0 and rowptr[n_nodes] == n_edges");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t* d_rowptr = nullptr;
    int32_t* d_col = nullptr;
    GIGL_CUDA(ctx, cudaMalloc(&d_rowptr, sizeof(int64_t) * (size_t)(n_nodes + 1)));
    cudaError_t e = cudaMalloc(&d_col, sizeof(int32_t) * (size_t)(n_edges > 0 ? n_edges : 1));
    if (e != cudaSuccess) {
        cudaFree(d_rowptr);
        return gigl_cuda_fail(ctx, e, "cudaMalloc(col)");
    }
    e = cudaMemcpyAsync(d_rowptr, rowptr, sizeof(int64_t) * (size_t)(n_nodes + 1), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n_edges > 0)
        e = cudaMemcpyAsync(d_col, col, sizeof(int32_t) * (size_t)n_edges, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d_rowptr);
        cudaFree(d_col);
        return gigl_cuda_fail(ctx, e, "graph upload");
    }
    int rc = graph_new(ctx, n_nodes, n_edges, d_rowptr, d_col, true, out);
    if (rc != GIGL_OK) {
        cudaFree(d_rowptr);
        cudaFree(d_col);
    }
    return rc;
}

int gigl_graph_wrap_dev(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int64_t* rowptr_dev,
                        const int32_t* col_dev, gigl_graph** out) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, out != nullptr, "null out pointer");
    *out = nullptr;
    GIGL_CHECK(ctx, n_nodes >= 0 && n_nodes <= 0x7fffffffLL, "n_nodes must be in [0, 2^31-1]");
    GIGL_CHECK(ctx, n_edges >= 0, "n_edges < 0");
    GIGL_CHECK(ctx, rowptr_dev != nullptr && (col_dev != nullptr || n_edges == 0), "null rowptr / col");
    return graph_new(ctx, n_nodes, n_edges, rowptr_dev, col_dev, false, out);
}

int gigl_graph_from_edges_dev(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src_dev,
                              const int32_t* dst_dev, int32_t is_graph_directed, int32_t by_source,
                              gigl_graph** out) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, out != nullptr, "null out pointer");
    *out = nullptr;
    GIGL_CHECK(ctx, n_nodes >= 0 && n_nodes <= 0x7fffffffLL, "n_nodes must be in [0, 2^31-1]");
    GIGL_CHECK(ctx, n_edges >= 0, "n_edges < 0");
    GIGL_CHECK(ctx, (src_dev && dst_dev) || n_edges == 0, "null edge arrays");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t* d_rowptr = nullptr;
    int32_t* d_col = nullptr;
    int64_t e_out = 0;
    int rc = graph_from_edges_build(ctx, n_nodes, n_edges, src_dev, dst_dev, is_graph_directed, by_source, &d_rowptr,
                                    &d_col, &e_out);
    if (rc != GIGL_OK) return rc;
    rc = graph_new(ctx, n_nodes, e_out, d_rowptr, d_col, true, out);
    if (rc != GIGL_OK) {
        cudaFree(d_rowptr);
        cudaFree(d_col);
    }
    return rc;
}

int gigl_graph_from_edges_host(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src,
                               const int32_t* dst, int32_t is_graph_directed, int32_t by_source,
                               gigl_graph** out) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, out != nullptr, "null out pointer");
    *out = nullptr;
    GIGL_CHECK(ctx, n_edges >= 0, "n_edges < 0");
    GIGL_CHECK(ctx, (src && dst) || n_edges == 0, "null edge arrays");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t* d_sd = nullptr;
    const size_t ne = (size_t)(n_edges > 0 ? n_edges : 1);
    GIGL_CUDA(ctx, cudaMalloc(&d_sd, sizeof(int32_t) * 2 * ne));
    cudaError_t e = cudaSuccess;
    if (n_edges > 0) {
        e = cudaMemcpyAsync(d_sd, src, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_sd + ne, dst, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e != cudaSuccess) {
        cudaFree(d_sd);
        return gigl_cuda_fail(ctx, e, "edge list upload");
    }
    int rc = gigl_graph_from_edges_dev(ctx, n_nodes, n_edges, d_sd, d_sd + ne, is_graph_directed, by_source, out);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_sd);
    return rc;
}

int gigl_edge_rows_host(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src, const int32_t* dst,
                        int32_t is_graph_directed, int32_t* edge_rows, int64_t rows_cap, int64_t* n_rows) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n_rows != nullptr, "null n_rows");
    *n_rows = 0;
    GIGL_CHECK(ctx, n_edges >= 0 && rows_cap >= 0, "negative size");
    GIGL_CHECK(ctx, n_edges <= 0x7fffffffLL, "edge record index must fit int32");
    GIGL_CHECK(ctx, (src && dst) || n_edges == 0, "null edge arrays");
    GIGL_CHECK(ctx, edge_rows || rows_cap == 0, "null edge_rows");
    if (n_edges == 0) return GIGL_OK;
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t *d_sd = nullptr, *d_rows = nullptr;
    const size_t ne = (size_t)n_edges;
    GIGL_CUDA(ctx, cudaMalloc(&d_sd, sizeof(int32_t) * 2 * ne));
    cudaError_t e = cudaMalloc(&d_rows, sizeof(int32_t) * (size_t)(rows_cap > 0 ? rows_cap : 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sd, src, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sd + ne, dst, sizeof(int32_t) * ne, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d_sd);
        cudaFree(d_rows);
        return gigl_cuda_fail(ctx, e, "edge list upload");
    }
    int rc = edge_rows_build(ctx, n_nodes, n_edges, d_sd, d_sd + ne, is_graph_directed, d_rows, rows_cap, n_rows);
    if (rc == GIGL_OK && *n_rows > 0) {
        e = cudaMemcpyAsync(edge_rows, d_rows, sizeof(int32_t) * (size_t)*n_rows, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = gigl_cuda_fail(ctx, e, "edge-row download");
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_sd);
    cudaFree(d_rows);
    return rc;
}

int gigl_graph_num_nodes(const gigl_graph* g, int64_t* n_nodes, int64_t* n_edges) {
    if (!g) return GIGL_E_INVALID;
    if (n_nodes) *n_nodes = g->n_nodes;
    if (n_edges) *n_edges = g->n_edges;
    return GIGL_OK;
}

int gigl_graph_device_ptrs(const gigl_graph* g, const int64_t** rowptr_dev, const int32_t** col_dev) {
    if (!g) return GIGL_E_INVALID;
    if (rowptr_dev) *rowptr_dev = g->rowptr;
    if (col_dev) *col_dev = g->col;
    return GIGL_OK;
}

void gigl_graph_destroy(gigl_graph* g) {
    if (!g) return;
    if (g->owned) {
        cudaSetDevice(g->ctx->device);
        cudaStreamSynchronize(g->ctx->stream);
        cudaFree((void*)g->rowptr);
        cudaFree((void*)g->col);
    }
    if (g->x_owned && g->x) {
        cudaSetDevice(g->ctx->device);
        cudaStreamSynchronize(g->ctx->stream);
        cudaFree((void*)g->x);
    }
    delete g;
}

// ---- sampling ------------------------------------------------------------------------------

int gigl_graph_set_hash_index(gigl_graph* g, int32_t enabled) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    g->hx_enabled = enabled != 0;
    return GIGL_OK;
}

int gigl_sample_khop_dev(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                         int32_t n_hops, int32_t base_seed, int32_t first_call_no, int32_t* const* nbr_dev,
                         int32_t* const* cnt_dev) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    GIGL_CUDA(g->ctx, cudaSetDevice(g->ctx->device));
    return khop_sample_launch(g, roots_dev, n_roots, fanouts, n_hops, base_seed, first_call_no, nbr_dev, cnt_dev);
}

// Samples hop by hop and forks the halo staging of every level as soon as it exists (batch_stage_*): the roots' and hop-h
// rows travel over NVLink while hop h + 1 is being sampled, the last hop's rows while the batch is collated.
static int khop_sample_staged(gigl_graph* g, gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts, int32_t n_hops,
                              int32_t base_seed, int32_t first_call_no, int32_t* const* nbr_dev, int32_t* const* cnt_dev) {
    int rc;
    if ((rc = batch_stage_begin(b, n_roots, fanouts, n_hops)) != GIGL_OK) return rc;
    gigl_stage_args stage{};
    if ((rc = batch_stage_args(b, &stage)) != GIGL_OK) return rc;
    // GIGL_STAGE_CLAIM=sampler: the sampling kernels claim the stage slots themselves (A/B; default = a claim pass per level
    // on the side stream)
    static const bool claim_in_sampler = getenv("GIGL_STAGE_CLAIM") && getenv("GIGL_STAGE_CLAIM")[0] == 's';
    if ((rc = batch_stage_level(b, roots_dev, n_roots, claim_in_sampler)) != GIGL_OK) return rc;
    int64_t width = n_roots;
    for (int h = 1; h <= n_hops; ++h) {
        if ((rc = khop_sample_launch(g, roots_dev, n_roots, fanouts, n_hops, base_seed, first_call_no, nbr_dev, cnt_dev, h, h,
                                     claim_in_sampler ? &stage : nullptr)) != GIGL_OK)
            return rc;
        width *= fanouts[h - 1];
        if ((rc = batch_stage_level(b, nbr_dev[h - 1], width, claim_in_sampler)) != GIGL_OK) return rc;
    }
    return batch_stage_end(b);
}

int gigl_sample_khop_staged_dev(gigl_graph* g, gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                                int32_t n_hops, int32_t base_seed, int32_t first_call_no, int32_t* const* nbr_dev,
                                int32_t* const* cnt_dev) {
    if (!g || !b) return gigl_fail(g ? g->ctx : nullptr, GIGL_E_INVALID, "null graph / batch");
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, batch_ctx(b) == ctx, "graph and batch must share one context");
    GIGL_CHECK(ctx, n_hops >= 1 && n_hops <= GIGL_MAX_HOPS && fanouts && nbr_dev && cnt_dev, "bad sampling arguments");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!batch_stages_early(b, g->x) || n_roots <= 0)  // nothing registered to stage from: plain sampling
        return khop_sample_launch(g, roots_dev, n_roots, fanouts, n_hops, base_seed, first_call_no, nbr_dev, cnt_dev);
    return khop_sample_staged(g, b, roots_dev, n_roots, fanouts, n_hops, base_seed, first_call_no, nbr_dev, cnt_dev);
}

int gigl_sample_op_dev(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, int32_t depth, const int32_t* chain_fanouts,
                       const int32_t* const* chain_nbr_dev, int32_t base_seed, int32_t call_no, int32_t* nbr_out_dev,
                       int32_t* cnt_out_dev) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, depth >= 1 && depth <= GIGL_MAX_HOPS && chain_fanouts, "depth must be in [1, 8]");
    GIGL_CHECK(ctx, (nbr_out_dev && cnt_out_dev) || n_roots == 0, "null output");
    GIGL_CHECK(ctx, depth == 1 || chain_nbr_dev, "null ancestor levels");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t* nbr[GIGL_MAX_HOPS];
    int32_t* cnt[GIGL_MAX_HOPS];
    for (int h = 0; h + 1 < depth; ++h) {
        GIGL_CHECK(ctx, chain_nbr_dev[h] || n_roots == 0, "null ancestor level");
        nbr[h] = const_cast<int32_t*>(chain_nbr_dev[h]);  // read only: hops below `depth` are not sampled by this call
        cnt[h] = cnt_out_dev;                              // never written
    }
    nbr[depth - 1] = nbr_out_dev;
    cnt[depth - 1] = cnt_out_dev;
    const int32_t first_call_no = (int32_t)((uint32_t)call_no - (uint32_t)(depth - 1));
    return khop_sample_launch(g, roots_dev, n_roots, chain_fanouts, depth, base_seed, first_call_no, nbr, cnt, depth, depth);
}

int gigl_sample_op_weighted_dev(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, int32_t depth, const int32_t* chain_fanouts,
                                const int32_t* const* chain_nbr_dev, const float* weights_dev, int32_t method, int32_t base_seed,
                                int32_t call_no, int32_t* nbr_out_dev, int32_t* cnt_out_dev) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, depth >= 1 && depth <= GIGL_MAX_HOPS && chain_fanouts, "depth must be in [1, 8]");
    GIGL_CHECK(ctx, (nbr_out_dev && cnt_out_dev) || n_roots == 0, "null output");
    GIGL_CHECK(ctx, depth == 1 || chain_nbr_dev, "null ancestor levels");
    GIGL_CHECK(ctx, method == GIGL_SAMPLE_TOP_K || method == GIGL_SAMPLE_RANDOM_WEIGHTED, "method must be GIGL_SAMPLE_TOP_K or GIGL_SAMPLE_RANDOM_WEIGHTED");
    GIGL_CHECK(ctx, weights_dev || g->n_edges == 0, "null edge weights");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t* nbr[GIGL_MAX_HOPS];
    int32_t* cnt[GIGL_MAX_HOPS];
    for (int h = 0; h + 1 < depth; ++h) {
        GIGL_CHECK(ctx, chain_nbr_dev[h] || n_roots == 0, "null ancestor level");
        nbr[h] = const_cast<int32_t*>(chain_nbr_dev[h]);  // read only
        cnt[h] = cnt_out_dev;                              // never written
    }
    nbr[depth - 1] = nbr_out_dev;
    cnt[depth - 1] = cnt_out_dev;
    const int32_t first_call_no = (int32_t)((uint32_t)call_no - (uint32_t)(depth - 1));
    static const float no_weights = 0.f;  // an edgeless graph: the kernel never dereferences it
    return khop_sample_launch(g, roots_dev, n_roots, chain_fanouts, depth, base_seed, first_call_no, nbr, cnt, depth, depth, nullptr,
                              weights_dev ? weights_dev : &no_weights, method);
}

int gigl_sample_op_host(gigl_graph* g, const int32_t* roots, int64_t n_roots, int32_t depth, const int32_t* chain_fanouts,
                        const int32_t* const* chain_nbr, int32_t base_seed, int32_t call_no, int32_t* nbr_out, int32_t* cnt_out) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, depth >= 1 && depth <= GIGL_MAX_HOPS && chain_fanouts, "depth must be in [1, 8]");
    GIGL_CHECK(ctx, n_roots >= 0 && (roots || n_roots == 0) && ((nbr_out && cnt_out) || n_roots == 0), "bad roots / output");
    if (n_roots == 0) return GIGL_OK;
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    // staging: roots | ancestor levels ... | cnt_out | nbr_out
    size_t width = 1, total = (size_t)n_roots;
    size_t off[GIGL_MAX_HOPS + 1];
    for (int h = 0; h < depth; ++h) {
        GIGL_CHECK(ctx, chain_fanouts[h] >= 1 && chain_fanouts[h] <= GIGL_MAX_FANOUT, "fanout must be in [1, 128]");
        if (h + 1 == depth) {
            off[depth] = total;  // cnt_out: one per parent slot
            total += (size_t)n_roots * width;
        }
        width *= (size_t)chain_fanouts[h];
        if ((double)n_roots * (double)width > 2147483647.0) return gigl_fail(ctx, GIGL_E_INVALID, "frontier exceeds 2^31-1 slots; split the roots");
        off[h] = total;
        total += (size_t)n_roots * width;
    }
    void* buf = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_IO0, sizeof(int32_t) * total, &buf);
    if (rc != GIGL_OK) return rc;
    int32_t* d = (int32_t*)buf;
    GIGL_CUDA(ctx, cudaMemcpyAsync(d, roots, sizeof(int32_t) * (size_t)n_roots, cudaMemcpyHostToDevice, ctx->stream));
    const int32_t* chain_dev[GIGL_MAX_HOPS];
    width = 1;
    for (int h = 0; h + 1 < depth; ++h) {
        width *= (size_t)chain_fanouts[h];
        GIGL_CHECK(ctx, chain_nbr && chain_nbr[h], "null ancestor level");
        GIGL_CUDA(ctx, cudaMemcpyAsync(d + off[h], chain_nbr[h], sizeof(int32_t) * (size_t)n_roots * width, cudaMemcpyHostToDevice, ctx->stream));
        chain_dev[h] = d + off[h];
    }
    const size_t parents = (size_t)n_roots * width;
    rc = gigl_sample_op_dev(g, d, n_roots, depth, chain_fanouts, chain_dev, base_seed, call_no, d + off[depth - 1], d + off[depth]);
    if (rc != GIGL_OK) return rc;
    GIGL_CUDA(ctx, cudaMemcpyAsync(cnt_out, d + off[depth], sizeof(int32_t) * parents, cudaMemcpyDeviceToHost, ctx->stream));
    GIGL_CUDA(ctx, cudaMemcpyAsync(nbr_out, d + off[depth - 1], sizeof(int32_t) * parents * (size_t)chain_fanouts[depth - 1],
                                   cudaMemcpyDeviceToHost, ctx->stream));
    return ctx_check_device_error(ctx);
}

int gigl_sample_khop_host(gigl_graph* g, const int32_t* roots, int64_t n_roots, const int32_t* fanouts,
                          int32_t n_hops, int32_t base_seed, int32_t first_call_no, int32_t* const* nbr,
                          int32_t* const* cnt) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, n_hops >= 1 && n_hops <= GIGL_MAX_HOPS, "n_hops must be in [1, 8]");
    GIGL_CHECK(ctx, n_roots >= 0, "n_roots < 0");
    GIGL_CHECK(ctx, fanouts && nbr && cnt, "null fanouts / output tables");
    GIGL_CHECK(ctx, roots || n_roots == 0, "null roots");
    for (int h = 0; h < n_hops; ++h) {
        GIGL_CHECK(ctx, fanouts[h] >= 1 && fanouts[h] <= GIGL_MAX_FANOUT, "fanout must be in [1, 128]");
        GIGL_CHECK(ctx, (nbr[h] && cnt[h]) || n_roots == 0, "null output level");
    }
    if (n_roots == 0) return GIGL_OK;
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    // one staging buffer: roots | cnt[0] nbr[0] | cnt[1] nbr[1] | ...  (int32 each)
    size_t total = (size_t)n_roots;
    size_t width = 1;
    size_t off_cnt[GIGL_MAX_HOPS], off_nbr[GIGL_MAX_HOPS];
    for (int h = 0; h < n_hops; ++h) {
        off_cnt[h] = total;
        total += (size_t)n_roots * width;
        width *= (size_t)fanouts[h];
        if ((double)n_roots * (double)width > 2147483647.0)
            return gigl_fail(ctx, GIGL_E_INVALID, "frontier exceeds 2^31-1 slots; split the roots");
        off_nbr[h] = total;
        total += (size_t)n_roots * width;
    }
    void* buf = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_IO0, sizeof(int32_t) * total, &buf);
    if (rc != GIGL_OK) return rc;
    int32_t* d = (int32_t*)buf;
    GIGL_CUDA(ctx, cudaMemcpyAsync(d, roots, sizeof(int32_t) * (size_t)n_roots, cudaMemcpyHostToDevice, ctx->stream));
    int32_t* nbr_dev[GIGL_MAX_HOPS];
    int32_t* cnt_dev[GIGL_MAX_HOPS];
    for (int h = 0; h < n_hops; ++h) {
        nbr_dev[h] = d + off_nbr[h];
        cnt_dev[h] = d + off_cnt[h];
    }
    rc = khop_sample_launch(g, d, n_roots, fanouts, n_hops, base_seed, first_call_no, nbr_dev, cnt_dev);
    if (rc != GIGL_OK) return rc;
    width = 1;
    for (int h = 0; h < n_hops; ++h) {
        GIGL_CUDA(ctx, cudaMemcpyAsync(cnt[h], cnt_dev[h], sizeof(int32_t) * (size_t)n_roots * width,
                                       cudaMemcpyDeviceToHost, ctx->stream));
        width *= (size_t)fanouts[h];
        GIGL_CUDA(ctx, cudaMemcpyAsync(nbr[h], nbr_dev[h], sizeof(int32_t) * (size_t)n_roots * width,
                                       cudaMemcpyDeviceToHost, ctx->stream));
    }
    return ctx_check_device_error(ctx);
}

int gigl_sample_positives_host(gigl_graph* g_out, const int32_t* srcs, int64_t n_srcs, int32_t num_pos,
                               int32_t base_seed, int32_t call_no, int32_t* pos, int32_t* pos_cnt) {
    int32_t* nbr[1] = {pos};
    int32_t* cnt[1] = {pos_cnt};
    return gigl_sample_khop_host(g_out, srcs, n_srcs, &num_pos, 1, base_seed, call_no, nbr, cnt);
}

// ---- aggregate -----------------------------------------------------------------------------

int gigl_csr_from_coo_dev(gigl_ctx* ctx, int64_t n, int64_t e, const int64_t* src_dev, const int64_t* dst_dev,
                          int64_t* rowptr_dev, int32_t* col_dev) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n >= 0 && n <= 0x7fffffffLL && e >= 0, "bad sizes");
    GIGL_CHECK(ctx, rowptr_dev != nullptr, "null rowptr");
    GIGL_CHECK(ctx, e == 0 || (src_dev && dst_dev && col_dev), "null edge arrays");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return csr_from_coo_launch(ctx, n, e, src_dev, dst_dev, rowptr_dev, col_dev);
}

int gigl_gather_mean_dev(gigl_ctx* ctx, int64_t n_rows_out, int32_t F, const int64_t* rowptr_dev,
                         const int32_t* col_dev, const float* x_dev, float* agg_dev) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n_rows_out >= 0 && F >= 1, "bad sizes");
    GIGL_CHECK(ctx, n_rows_out == 0 || (rowptr_dev && x_dev && agg_dev), "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return gather_mean_launch(ctx, n_rows_out, F, rowptr_dev, col_dev, x_dev, agg_dev);
}

int gigl_sage_conv_dev(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O,
                       const int64_t* rowptr_dev, const int32_t* col_dev, const float* x_dev, const float* Wl_dev,
                       const float* bl_dev, const float* Wr_dev, float* out_dev, int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n_rows_out == 0 || (rowptr_dev && x_dev && Wl_dev && Wr_dev && out_dev), "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return sage_conv_launch(ctx, n, n_rows_out, F, O, rowptr_dev, col_dev, x_dev, Wl_dev, bl_dev, Wr_dev, out_dev, relu);
}

int gigl_gcn_conv_dev(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr_dev,
                      const int32_t* col_dev, const float* x_dev, const float* W_dev, const float* b_dev,
                      float* out_dev, int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n == 0 || (rowptr_dev && x_dev && W_dev && out_dev), "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return gcn_conv_launch(ctx, n, F, O, rowptr_dev, col_dev, x_dev, W_dev, b_dev, out_dev, relu);
}

int gigl_sage_conv_host(gigl_ctx* ctx, int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* edge_index,
                        const float* x, const float* Wl, const float* bl, const float* Wr, float* out,
                        int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, Wl && Wr, "null weights");
    const float* w[3] = {Wl, bl, Wr};
    const size_t we[3] = {(size_t)O * F, (size_t)O, (size_t)O * F};
    return conv_host_common(ctx, n, e, F, O, edge_index, x, w, we, 3, out,
                            [&](const int64_t* rowptr, const int32_t* col, const float* xd, const float* const* wd, float* od) {
                                return sage_conv_launch(ctx, n, n, F, O, rowptr, col, xd, wd[0], wd[1], wd[2], od, relu);
                            });
}

int gigl_gcn_conv_host(gigl_ctx* ctx, int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* edge_index,
                       const float* x, const float* W, const float* b, float* out, int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, W != nullptr, "null weights");
    const float* w[2] = {W, b};
    const size_t we[2] = {(size_t)O * F, (size_t)O};
    return conv_host_common(ctx, n, e, F, O, edge_index, x, w, we, 2, out,
                            [&](const int64_t* rowptr, const int32_t* col, const float* xd, const float* const* wd, float* od) {
                                return gcn_conv_launch(ctx, n, F, O, rowptr, col, xd, wd[0], wd[1], od, relu);
                            });
}

// ---- features / model / batch -----------------------------------------------------------------

static void graph_drop_features(gigl_graph* g) {
    if (g->x_owned && g->x) cudaFree((void*)g->x);
    g->x = nullptr;
    g->F = 0;
    g->x_owned = false;
}

int gigl_graph_set_features_host(gigl_graph* g, const float* x, int32_t F) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, F >= 1 && (x != nullptr || g->n_nodes == 0), "bad feature table");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    GIGL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    graph_drop_features(g);
    float* d = nullptr;
    const size_t bytes = sizeof(float) * (size_t)(g->n_nodes > 0 ? g->n_nodes : 1) * (size_t)F;
    GIGL_CUDA(ctx, cudaMalloc(&d, bytes));
    cudaError_t e = cudaSuccess;
    if (g->n_nodes > 0) e = cudaMemcpyAsync(d, x, sizeof(float) * (size_t)g->n_nodes * F, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(d);
        return gigl_cuda_fail(ctx, e, "feature upload");
    }
    g->x = d;
    g->F = F;
    g->ldx = F;
    g->x_owned = true;
    return GIGL_OK;
}

int gigl_graph_set_features_dev(gigl_graph* g, const float* x_dev, int32_t F) { return gigl_graph_set_features_pitched_dev(g, x_dev, F, F); }

int gigl_graph_set_features_pitched_dev(gigl_graph* g, const float* x_dev, int32_t F, int64_t row_pitch) {
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    GIGL_CHECK(g->ctx, F >= 1 && x_dev != nullptr && row_pitch >= F, "bad feature table");
    graph_drop_features(g);
    g->x = x_dev;
    g->F = F;
    g->ldx = row_pitch;
    g->x_owned = false;
    return GIGL_OK;
}

int gigl_graph_features_dev(const gigl_graph* g, const float** x_dev, int32_t* F) {
    if (!g) return GIGL_E_INVALID;
    if (x_dev) *x_dev = g->x;
    if (F) *F = g->F;
    return GIGL_OK;
}

int gigl_linear_dev(gigl_ctx* ctx, int64_t M, int32_t N, int32_t K, const float* A_dev, int64_t lda, const float* W_dev,
                    int64_t ldw, const float* bias_dev, float* C_dev, int64_t ldc, int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, M >= 0 && N >= 1 && K >= 1 && lda >= K && ldw >= K && ldc >= N, "bad sizes");
    if (M == 0) return GIGL_OK;
    GIGL_CHECK(ctx, A_dev && W_dev && C_dev, "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    // split both operands into TF32 hi / lo halves (pitch padded to 4 floats), then 3xTF32 on tcgen05
    const int64_t ldp = ((int64_t)K + 3) & ~(int64_t)3;
    const size_t a_el = ((size_t)M * ldp + 63) & ~(size_t)63, w_el = ((size_t)N * ldp + 63) & ~(size_t)63;
    void* buf = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_AGG, sizeof(float) * 2 * (a_el + w_el), &buf);
    if (rc != GIGL_OK) return rc;
    float* a_hi = (float*)buf;
    float* a_lo = a_hi + a_el;
    float* w_hi = a_lo + a_el;
    float* w_lo = w_hi + w_el;
    if ((rc = split_tf32_launch(ctx, M, K, A_dev, lda, a_hi, a_lo, ldp)) != GIGL_OK) return rc;
    if ((rc = split_tf32_launch(ctx, N, K, W_dev, ldw, w_hi, w_lo, ldp)) != GIGL_OK) return rc;
    return linear_tc_launch(ctx, M, N, K, a_hi, a_lo, ldp, w_hi, w_lo, ldp, bias_dev, C_dev, ldc, relu);
}

int gigl_sage_conv_train_fwd_dev(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O, const int64_t* rowptr_dev,
                                 const int32_t* col_dev, const float* x_dev, const float* Wl_dev, const float* bl_dev,
                                 const float* Wr_dev, float* out_dev, float* saved_dev, int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n_rows_out == 0 || (rowptr_dev && x_dev && Wl_dev && Wr_dev && out_dev && saved_dev), "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return sage_conv_train_fwd_launch(ctx, n, n_rows_out, F, O, rowptr_dev, col_dev, x_dev, Wl_dev, bl_dev, Wr_dev, out_dev, saved_dev, relu);
}

int gigl_sage_conv_bwd_dev(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O, const int64_t* rowptr_dev,
                           const int64_t* t_rowptr_dev, const int32_t* t_col_dev, const float* saved_dev, const float* Wl_dev,
                           const float* Wr_dev, const float* out_dev, const float* grad_out_dev, float* grad_x_dev,
                           float* grad_Wl_dev, float* grad_bl_dev, float* grad_Wr_dev, int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, rowptr_dev && saved_dev && Wl_dev && Wr_dev && grad_out_dev, "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return sage_conv_bwd_launch(ctx, n, n_rows_out, F, O, rowptr_dev, t_rowptr_dev, t_col_dev, saved_dev, Wl_dev, Wr_dev, out_dev,
                                grad_out_dev, grad_x_dev, grad_Wl_dev, grad_bl_dev, grad_Wr_dev, relu);
}

int gigl_gcn_conv_bwd_dev(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr_dev, const int32_t* col_dev,
                          const int64_t* t_rowptr_dev, const int32_t* t_col_dev, const float* x_dev, const float* W_dev,
                          const float* out_dev, const float* grad_out_dev, float* grad_x_dev, float* grad_W_dev, float* grad_b_dev,
                          int32_t relu) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n == 0 || (rowptr_dev && x_dev && W_dev && grad_out_dev), "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return gcn_conv_bwd_launch(ctx, n, F, O, rowptr_dev, col_dev, t_rowptr_dev, t_col_dev, x_dev, W_dev, out_dev, grad_out_dev,
                               grad_x_dev, grad_W_dev, grad_b_dev, relu);
}

int gigl_linear_tn_dev(gigl_ctx* ctx, int64_t R, int32_t M, int32_t N, const float* G_dev, int64_t ldg, const float* A_dev,
                       int64_t lda, float* C_dev, int64_t ldc, int32_t accumulate) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, R >= 0 && M >= 1 && N >= 1 && ldg >= M && lda >= N && ldc >= N && C_dev, "bad sizes");
    GIGL_CHECK(ctx, R == 0 || (G_dev && A_dev), "null pointer");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return linear_tn_launch(ctx, R, M, N, G_dev, ldg, A_dev, lda, C_dev, ldc, accumulate);
}

int gigl_sage_model_create_host(gigl_ctx* ctx, int32_t n_layers, const int32_t* dims, const float* const* Wl,
                                const float* const* bl, const float* const* Wr, gigl_sage_model** out) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return sage_model_create(ctx, n_layers, dims, Wl, bl, Wr, 0, out);
}

int gigl_sage_model_create_dev(gigl_ctx* ctx, int32_t n_layers, const int32_t* dims, const float* const* Wl_dev,
                               const float* const* bl_dev, const float* const* Wr_dev, gigl_sage_model** out) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return sage_model_create(ctx, n_layers, dims, Wl_dev, bl_dev, Wr_dev, 1, out);
}

void gigl_sage_model_destroy(gigl_sage_model* m) { sage_model_destroy(m); }

int gigl_batch_create(gigl_ctx* ctx, int64_t n_graph_nodes, gigl_batch** out) {
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, out != nullptr && n_graph_nodes >= 0 && n_graph_nodes <= 0x7fffffffLL, "bad arguments");
    *out = nullptr;
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return batch_create(ctx, n_graph_nodes, out);
}

void gigl_batch_destroy(gigl_batch* b) { batch_destroy(b); }

int gigl_batch_collate_dev(gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                           int32_t n_hops, const int32_t* const* nbr_dev, int32_t n_layers, int64_t* level_sizes,
                           int64_t* n_edges) {
    if (!b) return gigl_fail(nullptr, GIGL_E_INVALID, "null batch");
    gigl_ctx* ctx = batch_ctx(b);
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = batch_collate(b, roots_dev, n_roots, fanouts, n_hops, nbr_dev, n_layers, level_sizes, n_edges);
    if (rc != GIGL_OK) return rc;
    return ctx_check_device_error(ctx);
}

int gigl_batch_finalize_nodes(gigl_batch* b, int64_t* n_nodes, int64_t* n_edges) {
    if (!b) return gigl_fail(nullptr, GIGL_E_INVALID, "null batch");
    GIGL_CUDA(batch_ctx(b), cudaSetDevice(batch_ctx(b)->device));
    return batch_finalize_nodes(b, n_nodes, n_edges);
}

int gigl_batch_set_halo_staging(gigl_batch* b, int32_t enabled) {
    if (!b) return gigl_fail(nullptr, GIGL_E_INVALID, "null batch");
    batch_set_halo_staging(b, enabled != 0);
    return GIGL_OK;
}

int gigl_batch_set_halo_table_dev(gigl_batch* b, const float* x_dev, int32_t F, int64_t ldx) {
    if (!b) return gigl_fail(nullptr, GIGL_E_INVALID, "null batch");
    GIGL_CUDA(batch_ctx(b), cudaSetDevice(batch_ctx(b)->device));
    return batch_set_halo_table(b, x_dev, F, ldx);
}

int gigl_batch_set_hot_rows_dev(gigl_batch* b, const int32_t* hot_slot_dev, const float* hot_dev, int32_t F, int64_t ld) {
    if (!b) return gigl_fail(nullptr, GIGL_E_INVALID, "null batch");
    return batch_set_hot_rows(b, hot_slot_dev, hot_dev, F, ld);
}

int gigl_batch_export_dev(gigl_batch* b, int32_t* node_ids_dev, int64_t* edge_index_dev) {
    if (!b) return gigl_fail(nullptr, GIGL_E_INVALID, "null batch");
    GIGL_CUDA(batch_ctx(b), cudaSetDevice(batch_ctx(b)->device));
    return batch_export(b, node_ids_dev, edge_index_dev);
}

int gigl_batch_sage_forward_dev(gigl_batch* b, const gigl_sage_model* m, const float* x_dev, int64_t ldx,
                                float* out_dev) {
    if (!b) return gigl_fail(nullptr, GIGL_E_INVALID, "null batch");
    gigl_ctx* ctx = batch_ctx(b);
    GIGL_CHECK(ctx, m != nullptr, "null model");
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    return batch_sage_forward(b, m, x_dev, ldx, out_dev);
}

// Shared body of the two host entry points.  packed == nullptr: padded-tree index sets into nbr / cnt (either may be
// NULL).  packed != nullptr: one-byte counts into cnt_u8[h] and the filled slots into packed (tree_pack.cu).
static int infer_khop_sage_host_impl(gigl_graph* g, gigl_batch* b, const gigl_sage_model* m, const int32_t* roots, int64_t n_roots,
                                     const int32_t* fanouts, int32_t n_hops, int32_t base_seed, int32_t first_call_no, float* out,
                                     int32_t* const* nbr, int32_t* const* cnt, uint8_t* const* cnt_u8, int32_t* packed,
                                     int64_t packed_cap, int64_t* n_packed, int32_t id_bits = 0) {
    // id_bits > 0: `packed` is a bit stream of id_bits-bit ids in 32-bit words (packed_cap counts words)
    if (!g) return gigl_fail(nullptr, GIGL_E_INVALID, "null graph");
    gigl_ctx* ctx = g->ctx;
    GIGL_CHECK(ctx, b != nullptr && m != nullptr && batch_ctx(b) == ctx, "batch / model must belong to the graph's ctx");
    GIGL_CHECK(ctx, g->x != nullptr, "the graph holds no features (gigl_graph_set_features_*)");
    GIGL_CHECK(ctx, n_hops >= 1 && n_hops <= GIGL_MAX_HOPS && fanouts, "n_hops must be in [1, 8]");
    GIGL_CHECK(ctx, n_roots >= 0 && (roots || n_roots == 0) && (out || n_roots == 0), "bad roots / out");
    const bool packing = packed != nullptr;
    GIGL_CHECK(ctx, !packing || (cnt_u8 != nullptr && n_packed != nullptr && packed_cap >= 0), "packed output needs cnt_u8, n_packed");
    int32_t n_layers = 0, dims[GIGL_MAX_HOPS + 1];
    sage_model_dims(m, &n_layers, dims);
    GIGL_CHECK(ctx, dims[0] == g->F, "model input width != feature width");
    if (n_packed) *n_packed = 0;
    if (n_roots == 0) return GIGL_OK;
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    // device layout: roots | cnt[0] .. cnt[H-1] (back to back: tree_pack scans them as one array) | nbr[0] .. nbr[H-1]
    size_t total = (size_t)n_roots, width = 1, n_parents = 0;
    size_t off_cnt[GIGL_MAX_HOPS], off_nbr[GIGL_MAX_HOPS];
    for (int h = 0; h < n_hops; ++h) {
        GIGL_CHECK(ctx, fanouts[h] >= 1 && fanouts[h] <= GIGL_MAX_FANOUT, "fanout must be in [1, 128]");
        off_cnt[h] = total;
        total += (size_t)n_roots * width;
        n_parents += (size_t)n_roots * width;
        width *= (size_t)fanouts[h];
        if ((double)n_roots * (double)width > 2147483647.0)
            return gigl_fail(ctx, GIGL_E_INVALID, "frontier exceeds 2^31-1 slots; split the roots");
    }
    width = 1;
    for (int h = 0; h < n_hops; ++h) {
        width *= (size_t)fanouts[h];
        off_nbr[h] = total;
        total += (size_t)n_roots * width;
    }
    void *buf = nullptr, *pout = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_IO0, sizeof(int32_t) * (total + 1), &buf);
    if (rc != GIGL_OK) return rc;
    const int O = dims[n_layers];
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_IO1, sizeof(float) * (size_t)n_roots * O, &pout)) != GIGL_OK) return rc;
    int32_t* d = (int32_t*)buf;
    GIGL_CUDA(ctx, cudaMemcpyAsync(d, roots, sizeof(int32_t) * (size_t)n_roots, cudaMemcpyHostToDevice, ctx->stream));
    int32_t* nbr_dev[GIGL_MAX_HOPS];
    int32_t* cnt_dev[GIGL_MAX_HOPS];
    for (int h = 0; h < n_hops; ++h) {
        nbr_dev[h] = d + off_nbr[h];
        cnt_dev[h] = d + off_cnt[h];
    }
    // The index sets (59.5 MB for B = 65536, [15, 10]; ~40 MB packed) leave on a second stream while the aggregate runs.
    // Measured on B200: a device-to-host copy that overlaps the SAMPLER or the collation sort slows them by about the
    // copy's own duration (both live on L2-resident tables - the 56 MB hash-key table, the radix-sort buffers - and the
    // copy streams 48 MB through the same L2), while the gather / projection kernels stream gigabytes through L2 anyway.
    // So the copies are released after collation and ride under the aggregate (0.9 ms of PCIe under 1.4 ms of kernels).
    const bool copying = packing || (nbr && cnt);
    static const bool dbg = getenv("GIGL_DEBUG_TIMELINE") != nullptr;
    cudaEvent_t ev[8] = {};
    if (dbg) {
        for (auto& e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], ctx->stream);
    }
    if (copying && !ctx->copy_stream) {
        GIGL_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        GIGL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_ready, cudaEventDisableTiming));
    }
    if (packing && !ctx->pack_ready) {
        GIGL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pack_ready, cudaEventDisableTiming));
        GIGL_CUDA(ctx, cudaMallocHost(&ctx->h_pack_total, sizeof(int32_t)));
    }
    if (batch_stages_early(b, g->x) && n_roots > 0)
        rc = khop_sample_staged(g, b, d, n_roots, fanouts, n_hops, base_seed, first_call_no, nbr_dev, cnt_dev);
    else
        rc = khop_sample_launch(g, d, n_roots, fanouts, n_hops, base_seed, first_call_no, nbr_dev, cnt_dev);
    if (rc != GIGL_OK) return rc;
    if (dbg) cudaEventRecord(ev[2], ctx->stream);
    int32_t* goff_dev = nullptr;
    uint8_t* u8_dev = nullptr;
    int32_t* packed_dev = nullptr;
    uint32_t* words_dev = nullptr;
    if (packing) {
        // the pack kernels (one scan + ~70 MB of traffic, tens of microseconds) run on the copy stream beside the collation
        GIGL_CUDA(ctx, cudaEventRecord(ctx->copy_ready, ctx->stream));
        GIGL_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ready, 0));
        if ((rc = tree_pack_launch(ctx, ctx->copy_stream, n_roots, fanouts, n_hops, nbr_dev, cnt_dev[0], GIGL_SLOT_IO2, &goff_dev, &u8_dev,
                                   &packed_dev, id_bits, &words_dev)) != GIGL_OK)
            return rc;
        GIGL_CUDA(ctx, cudaMemcpyAsync(ctx->h_pack_total, goff_dev + n_parents, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
        GIGL_CUDA(ctx, cudaEventRecord(ctx->pack_ready, ctx->copy_stream));
    }
    if ((rc = batch_collate(b, d, n_roots, fanouts, n_hops, nbr_dev, n_layers, nullptr, nullptr)) != GIGL_OK) {
        if (packing) cudaStreamSynchronize(ctx->copy_stream);
        return rc;
    }
    if (dbg) cudaEventRecord(ev[4], ctx->stream);
    if (copying) {
        GIGL_CUDA(ctx, cudaEventRecord(ctx->copy_ready, ctx->stream));
        GIGL_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_ready, 0));
        if (dbg) cudaEventRecord(ev[1], ctx->copy_stream);
        if (packing) {
            GIGL_CUDA(ctx, cudaEventSynchronize(ctx->pack_ready));  // long done: the collation synchronised the main stream twice
            const int64_t filled = *ctx->h_pack_total;
            *n_packed = filled;
            const int64_t n_words = id_bits > 0 ? (filled * id_bits + 31) / 32 : filled;
            if (n_words > packed_cap) return gigl_fail(ctx, GIGL_E_INVALID, "packed buffer too small for the sampled index sets");
            size_t p0 = 0;
            width = 1;
            for (int h = 0; h < n_hops; ++h) {
                const size_t parents = (size_t)n_roots * width;
                if (cnt_u8[h]) GIGL_CUDA(ctx, cudaMemcpyAsync(cnt_u8[h], u8_dev + p0, parents, cudaMemcpyDeviceToHost, ctx->copy_stream));
                p0 += parents;
                width *= (size_t)fanouts[h];
            }
            if (filled > 0)
                GIGL_CUDA(ctx, cudaMemcpyAsync(packed, id_bits > 0 ? (const void*)words_dev : (const void*)packed_dev, sizeof(int32_t) * (size_t)n_words,
                                               cudaMemcpyDeviceToHost, ctx->copy_stream));
        } else {
            width = 1;
            for (int h = 0; h < n_hops; ++h) {
                if (cnt[h]) GIGL_CUDA(ctx, cudaMemcpyAsync(cnt[h], cnt_dev[h], sizeof(int32_t) * (size_t)n_roots * width, cudaMemcpyDeviceToHost, ctx->copy_stream));
                width *= (size_t)fanouts[h];
                if (nbr[h]) GIGL_CUDA(ctx, cudaMemcpyAsync(nbr[h], nbr_dev[h], sizeof(int32_t) * (size_t)n_roots * width, cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
        }
        if (dbg) cudaEventRecord(ev[3], ctx->copy_stream);
    }
    rc = batch_sage_forward(b, m, g->x, g->ldx, (float*)pout);
    if (rc != GIGL_OK) {
        if (copying) cudaStreamSynchronize(ctx->copy_stream);  // nothing may still be writing the caller's buffers
        return rc;
    }
    if (dbg) cudaEventRecord(ev[5], ctx->stream);
    GIGL_CUDA(ctx, cudaMemcpyAsync(out, pout, sizeof(float) * (size_t)n_roots * O, cudaMemcpyDeviceToHost, ctx->stream));
    if (dbg) cudaEventRecord(ev[6], ctx->stream);
    if (copying) {
        // join: the ctx stream (and whoever times it) is not done before the index sets have landed
        GIGL_CUDA(ctx, cudaEventRecord(ctx->copy_ready, ctx->copy_stream));
        GIGL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_ready, 0));
    }
    if (dbg) {
        cudaEventRecord(ev[7], ctx->stream);
        cudaEventSynchronize(ev[7]);
        float t[8] = {};
        for (int i = 1; i < 8; ++i)
            if (copying || (i != 1 && i != 3)) cudaEventElapsedTime(&t[i], ev[0], ev[i]);
        fprintf(stderr, "[timeline ms] first copy start %.3f | sampled %.3f | copies done %.3f | collated %.3f | forward %.3f | out copied %.3f | joined %.3f\n",
                t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
        for (auto& e : ev) cudaEventDestroy(e);
    }
    return ctx_check_device_error(ctx);
}

int gigl_infer_khop_sage_host(gigl_graph* g, gigl_batch* b, const gigl_sage_model* m, const int32_t* roots,
                              int64_t n_roots, const int32_t* fanouts, int32_t n_hops, int32_t base_seed,
                              int32_t first_call_no, float* out, int32_t* const* nbr, int32_t* const* cnt) {
    return infer_khop_sage_host_impl(g, b, m, roots, n_roots, fanouts, n_hops, base_seed, first_call_no, out, nbr, cnt, nullptr, nullptr, 0,
                                     nullptr);
}

int gigl_infer_khop_sage_packed_host(gigl_graph* g, gigl_batch* b, const gigl_sage_model* m, const int32_t* roots, int64_t n_roots,
                                     const int32_t* fanouts, int32_t n_hops, int32_t base_seed, int32_t first_call_no, float* out,
                                     uint8_t* const* cnt_u8, int32_t* packed, int64_t packed_cap, int64_t* n_packed) {
    if (g && !packed) return gigl_fail(g->ctx, GIGL_E_INVALID, "null packed buffer");
    return infer_khop_sage_host_impl(g, b, m, roots, n_roots, fanouts, n_hops, base_seed, first_call_no, out, nullptr, nullptr, cnt_u8, packed,
                                     packed_cap, n_packed);
}

static int32_t id_bits_of(int64_t n_nodes) {
    int32_t bits = 1;
    while (bits < 31 && (1LL << bits) < n_nodes) ++bits;
    return bits;
}

int gigl_infer_khop_sage_bitpacked_host(gigl_graph* g, gigl_batch* b, const gigl_sage_model* m, const int32_t* roots, int64_t n_roots,
                                        const int32_t* fanouts, int32_t n_hops, int32_t base_seed, int32_t first_call_no, float* out,
                                        uint8_t* const* cnt_u8, uint32_t* words, int64_t words_cap, int64_t* n_packed, int32_t* id_bits) {
    if (g && (!words || !id_bits)) return gigl_fail(g->ctx, GIGL_E_INVALID, "null words / id_bits");
    if (g) *id_bits = id_bits_of(g->n_nodes);
    return infer_khop_sage_host_impl(g, b, m, roots, n_roots, fanouts, n_hops, base_seed, first_call_no, out, nullptr, nullptr, cnt_u8,
                                     reinterpret_cast<int32_t*>(words), words_cap, n_packed, g ? *id_bits : 0);
}

int gigl_unpack_bits_host(const uint32_t* words, int64_t n, int32_t bits, int32_t* out) {
    if (n < 0 || bits < 1 || bits > 31 || ((!words || !out) && n > 0)) return GIGL_E_INVALID;
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t bit = i * bits;
        const int64_t w = bit >> 5;
        const int s = (int)(bit & 31);
        uint32_t v = words[w] >> s;
        if (s + bits > 32) v |= words[w + 1] << (32 - s);
        out[i] = (int32_t)(v & mask);
    }
    return GIGL_OK;
}

}  // extern "C"
