// common.cuh - context object, error plumbing and launch helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/gigl_b200.h"

// scratch slot ids: a launch sequence that needs several live buffers uses distinct slots
enum {
    GIGL_SLOT_WORK = 0,   // sampler worklists / CUB temp storage
    GIGL_SLOT_AGG = 1,    // aggregate intermediates (agg rows, x', dinv)
    GIGL_SLOT_SORT = 2,   // COO->CSR sort keys / values
    GIGL_SLOT_IO0 = 3,    // host-entry-point staging: inputs
    GIGL_SLOT_IO1 = 4,    // host-entry-point staging: outputs
    GIGL_SLOT_IO2 = 5,    // host-entry-point staging: csr
    GIGL_SLOT_IO3 = 6,
    GIGL_SLOT_IO4 = 7,
    GIGL_SLOT_SAVE = 8,   // [mean | self] of an inference-only layer / operand halves of the GCN projection
    GIGL_SCRATCH_SLOTS = 9
};

// timing tags: device time per phase, measured with CUDA events on the ctx stream when enabled
enum {
    GIGL_T_SAMPLE = 0,     // k-hop sampling kernels (all hops)
    GIGL_T_COLLATE_KEYS,   // tree slots -> edge keys
    GIGL_T_COLLATE_SORT,   // radix sort of the keys
    GIGL_T_COLLATE_MAPS,   // row bounds + level / local-id assignment (+ cleanup of the previous batch)
    GIGL_T_GATHER_L1,      // layer-1 gather (reads the graph-wide feature table): main + split-row launches
    GIGL_T_GATHER_DEEP,    // gathers of layers >= 2
    GIGL_T_GEMM_L1,        // layer-1 projection
    GIGL_T_GEMM_DEEP,      // projections of layers >= 2
    GIGL_T_GATHER_FULL,    // full-graph gather (gigl_sage_conv_dev / gigl_gather_mean_dev)
    GIGL_T_GEMM_FULL,
    GIGL_T_HALO_STAGE,     // per-batch copy of the unique nodes' feature rows (sharded feature table: the NVLink halo)
    GIGL_T_HALO_WAIT,      // early staging: what the ctx stream still waits for the side stream's copy before layer 1
    GIGL_T_COUNT
};

struct gigl_timer_pair {
    cudaEvent_t a, b;
    int tag;
};

// Hash ladder of the sampler (khop_sample.cu): the permutation key of hash input x is a fixed sequence H(x), so the
// "f smallest keys of a window" query is indexed once per context, for every graph sampled through it.
//   level 0: tk0[x] = top 32 bits of ordered_key(x), every x in [0, limit);
//   level j >= 1: the inputs whose key has j leading zero bits (a 2^-j sample of the inputs), as entries
//   (key >> (32 - j)) << 32 | x grouped by block x >> j; bs[j][b] = first entry of block b.
constexpr int GIGL_LAD_MAX_LEVELS = 25;
struct gigl_ladder {
    uint32_t* tk0 = nullptr;
    uint32_t* bs[GIGL_LAD_MAX_LEVELS + 1] = {};
    uint64_t* ent[GIGL_LAD_MAX_LEVELS + 1] = {};
    uint64_t limit = 0;  // inputs [0, limit) are covered, limit <= 2^31
    int levels = 0;
    void* blob_a = nullptr;  // tk0 + block starts
    void* blob_b = nullptr;  // entries
};

struct gigl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // second stream + event for device->host copies that overlap the kernels of the same call (host entry points)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ready = nullptr;
    cudaEvent_t pack_ready = nullptr;   // the packed index sets (tree_pack.cu) are complete on copy_stream
    int32_t* h_pack_total = nullptr;    // pinned: filled slots of the packed tree
    int sm_count = 148;
    int64_t launches = 0;
    std::string err;
    // device-side error word written by kernels (0 = ok, else a GIGL_E_* code) + its pinned host mirror
    int32_t* d_err = nullptr;
    int32_t* h_err = nullptr;
    // growable device scratch slots (worklists, partial buffers, staging); never shrink
    void* scratch[GIGL_SCRATCH_SLOTS] = {};
    size_t scratch_bytes[GIGL_SCRATCH_SLOTS] = {};
    // phase timing (off by default)
    bool timing = false;
    gigl_timer_pair* t_pool = nullptr;
    int t_cap = 0, t_used = 0;
    double t_ms[GIGL_T_COUNT] = {};
    int64_t t_n[GIGL_T_COUNT] = {};
    gigl_ladder lad;  // sampler's hash ladder, built lazily on the first sampling call
    // host waits inside the entry points: GIGL_SYNC=block sleeps on a blocking-sync event instead of spinning in
    // cudaStreamSynchronize - for boxes with fewer host cores than caller threads (8 GPUs x batches in flight)
    bool block_sync = false;
    cudaEvent_t ev_block = nullptr;
};

// The host waits for everything queued on `s` (a stream of this context).
static inline cudaError_t gigl_host_wait(gigl_ctx* ctx, cudaStream_t s) {
    if (!ctx->block_sync) return cudaStreamSynchronize(s);
    const cudaError_t e = cudaEventRecord(ctx->ev_block, s);
    return e != cudaSuccess ? e : cudaEventSynchronize(ctx->ev_block);
}

// Begin / end of a timed phase on the ctx stream (no-ops unless timing is enabled).
int gigl_timer_begin(gigl_ctx* ctx, int tag);
void gigl_timer_end(gigl_ctx* ctx, int handle);
int gigl_timer_begin_on(gigl_ctx* ctx, int tag, cudaStream_t stream);  // the same on a side stream of the context
void gigl_timer_end_on(gigl_ctx* ctx, int handle, cudaStream_t stream);
struct gigl_timed {  // RAII helper
    gigl_ctx* ctx;
    int h;
    gigl_timed(gigl_ctx* c, int tag) : ctx(c), h(gigl_timer_begin(c, tag)) {}
    ~gigl_timed() { gigl_timer_end(ctx, h); }
};

struct gigl_graph {
    gigl_ctx* ctx = nullptr;
    int64_t n_nodes = 0;
    int64_t n_edges = 0;
    const int64_t* rowptr = nullptr;  // device
    const int32_t* col = nullptr;     // device
    bool owned = false;
    const float* x = nullptr;  // device feature table [n_nodes, F] (optional), row pitch ldx floats
    int32_t F = 0;
    int64_t ldx = 0;
    bool x_owned = false;
    bool hx_enabled = true;  // sample through the context's hash ladder (gigl_graph_set_hash_index)
};

int gigl_fail(gigl_ctx* ctx, int code, const std::string& msg);
int gigl_cuda_fail(gigl_ctx* ctx, cudaError_t e, const char* what);
// Ensures ctx->scratch[slot] holds >= bytes; returns GIGL_OK or an error code.
int gigl_scratch(gigl_ctx* ctx, int slot, size_t bytes, void** out);

#define GIGL_CUDA(ctx, call)                                        \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return gigl_cuda_fail(ctx, e__, #call); \
    } while (0)

#define GIGL_CHECK(ctx, cond, msg)                               \
    do {                                                         \
        if (!(cond)) return gigl_fail(ctx, GIGL_E_INVALID, msg); \
    } while (0)

// Checks the launch and bumps the ctx launch counter.
#define GIGL_LAUNCHED(ctx)                                                  \
    do {                                                                    \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) return gigl_cuda_fail(ctx, e__, "kernel launch"); \
        (ctx)->launches++;                                                  \
    } while (0)

#ifdef __CUDACC__
// 3xTF32 operand split: hi = v rounded to nearest TF32 (10-bit mantissa), lo = (v - hi) rounded to TF32 as well, so
// the tensor core's own truncation of its inputs is a no-op and every error term (lo rounding, the dropped lo*lo
// product) is <= 2^-22 relative and unbiased - truncating instead gives a one-sided 2^-20 error that adds up linearly
// along long reductions (measured on the backward pass of hub rows).
__device__ __forceinline__ float gigl_tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void gigl_split_tf32(float v, float& hi, float& lo) {
    hi = gigl_tf32_rna(v);
    lo = gigl_tf32_rna(v - hi);
}
#endif

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- entry points implemented in the .cu files, called from capi.cu ---------------------
// Stage-slot claims done by the sampling kernels themselves (sharded feature table, batch_stage_*): every vertex a hop writes
// into the tree is claimed in `slot` (dense map, kStageAbsent = free) and appended to `list`; *ctr = claimed so far.
struct gigl_stage_args {
    int32_t* slot;
    int32_t* list;
    int32_t* ctr;
};
constexpr int32_t kStageAbsent = 0x7fffffff;   // == kLidAbsent of batch_collate.cu
constexpr int32_t kStagePending = 0x7ffffffe;  // claimed, slot number not written yet
int khop_sample_launch(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                       int32_t n_hops, int32_t base_seed, int32_t first_call_no, int32_t* const* nbr_dev,
                       int32_t* const* cnt_dev, int32_t hop_first = 1, int32_t hop_last = 0,  // hops [hop_first, hop_last], 0 = n_hops
                       const gigl_stage_args* stage = nullptr,
                       const float* weights_dev = nullptr, int32_t method = 0);  // method: GIGL_SAMPLE_* (weights by CSR position)
int csr_from_coo_launch(gigl_ctx* ctx, int64_t n, int64_t e, const int64_t* src, const int64_t* dst,
                        int64_t* rowptr, int32_t* col);
int graph_from_edges_build(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src_dev,
                           const int32_t* dst_dev, int32_t directed, int32_t by_source, int64_t** rowptr_dev,
                           int32_t** col_dev, int64_t* n_edges_out);
int edge_rows_build(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src_dev, const int32_t* dst_dev,
                    int32_t directed, int32_t* rows_dev, int64_t rows_cap, int64_t* n_rows_out);
int gather_mean_launch(gigl_ctx* ctx, int64_t n_rows, int32_t F, const int64_t* rowptr, const int32_t* col,
                       const float* x, float* agg);
int sage_conv_launch(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O, const int64_t* rowptr,
                     const int32_t* col, const float* x, const float* Wl, const float* bl, const float* Wr,
                     float* out, int32_t relu);
int gcn_conv_launch(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr, const int32_t* col,
                    const float* x, const float* W, const float* b, float* out, int32_t relu);

// tree_pack.cu: padded tree -> goff (exclusive scan of the counts; goff[n_parents] = filled slots), one-byte counts, packed children
int tree_pack_launch(gigl_ctx* ctx, cudaStream_t st, int64_t n_roots, const int32_t* fanouts, int32_t n_hops,
                     const int32_t* const* nbr_dev, const int32_t* cnt_all_dev, int slot, int32_t** goff_dev, uint8_t** cnt_u8_dev,
                     int32_t** packed_dev, int id_bits = 0, uint32_t** words_dev = nullptr);

// batch_collate.cu
struct gigl_batch;
struct gigl_sage_model;
int batch_create(gigl_ctx* ctx, int64_t n_graph_nodes, gigl_batch** out);
void batch_destroy(gigl_batch* b);
int batch_collate(gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts, int32_t n_hops,
                  const int32_t* const* nbr_dev, int32_t n_layers, int64_t* level_sizes_host, int64_t* n_edges_host);
int batch_finalize_nodes(gigl_batch* b, int64_t* n_nodes, int64_t* n_edges);
void batch_set_halo_staging(gigl_batch* b, bool enabled);
int batch_set_halo_table(gigl_batch* b, const float* x_dev, int32_t F, int64_t ldx);
bool batch_stages_early(const gigl_batch* b, const float* x_dev);
int batch_stage_begin(gigl_batch* b, int64_t n_roots, const int32_t* fanouts, int32_t n_hops);
int batch_stage_args(gigl_batch* b, gigl_stage_args* out);
int batch_stage_level(gigl_batch* b, const int32_t* ids_dev, int64_t n, bool claimed);  // claimed: by the kernel that wrote ids
int batch_stage_end(gigl_batch* b);
int batch_set_hot_rows(gigl_batch* b, const int32_t* hot_slot_dev, const float* hot_dev, int32_t F, int64_t ld);
int batch_export(gigl_batch* b, int32_t* node_ids_dev, int64_t* edge_index_dev);
int batch_sage_forward(gigl_batch* b, const gigl_sage_model* m, const float* x_dev, int64_t ldx, float* out_dev);
gigl_ctx* batch_ctx(gigl_batch* b);
int sage_model_create(gigl_ctx* ctx, int32_t n_layers, const int32_t* dims, const float* const* Wl, const float* const* bl,
                      const float* const* Wr, int weights_on_device, gigl_sage_model** out);
void sage_model_destroy(gigl_sage_model* m);
int sage_model_dims(const gigl_sage_model* m, int32_t* n_layers, int32_t* dims);

// gemm_tcgen05.cu: C[M, N] = (A_hi + A_lo)[M, K] @ (W_hi + W_lo)[N, K]^T + bias as 3xTF32 on tcgen05 (TMA-fed, TMEM accumulators)
int linear_tc_launch(gigl_ctx* ctx, int64_t M, int N, int K, const float* A_hi, const float* A_lo, int64_t lda,
                     const float* W_hi, const float* W_lo, int64_t ldw, const float* bias, float* C, int64_t ldc, int relu);
int linear_tc_launch_ex(gigl_ctx* ctx, int64_t M, const int32_t* m_dev, int N, int K, const float* A_hi, const float* A_lo, int64_t lda,
                        const float* W_hi, const float* W_lo, int64_t ldw, const float* bias, float* C, int64_t ldc, int relu);
int split_tf32_launch(gigl_ctx* ctx, int64_t rows, int cols, const float* x, int64_t ldx, float* hi, float* lo, int64_t ldo);
// gemm_tn_tcgen05.cu: C = G^T A (weight gradients; split-K over the rows, deterministic), column sums (bias gradient)
int linear_tn_tc_launch(gigl_ctx* ctx, int64_t R, int M, int N, const float* G_hi, const float* G_lo, int64_t ldg, const float* A_hi,
                        const float* A_lo, int64_t lda, float* C0, int64_t ldc0, int n0_valid, int n_split, float* C1, int64_t ldc1,
                        int n1_valid, int accumulate);
int colsum_launch(gigl_ctx* ctx, int64_t R, int M, const float* G, int64_t ldg, float* out, int accumulate);
// sage_aggregate.cu: training forms (forward that keeps [mean | self], backward)
int sage_conv_train_fwd_launch(gigl_ctx* ctx, int64_t n, int64_t m, int32_t F, int32_t O, const int64_t* rowptr, const int32_t* col,
                               const float* x, const float* Wl, const float* bl, const float* Wr, float* out, float* A_save, int32_t relu);
int sage_conv_bwd_launch(gigl_ctx* ctx, int64_t n, int64_t m, int32_t F, int32_t O, const int64_t* rowptr, const int64_t* t_rowptr,
                         const int32_t* t_col, const float* A_save, const float* Wl, const float* Wr, const float* out,
                         const float* grad_out, float* grad_x, float* grad_Wl, float* grad_bl, float* grad_Wr, int32_t relu);
int gcn_conv_bwd_launch(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr, const int32_t* col, const int64_t* t_rowptr,
                        const int32_t* t_col, const float* x, const float* W, const float* out, const float* grad_out, float* grad_x,
                        float* grad_W, float* grad_b, int32_t relu);
int linear_tn_launch(gigl_ctx* ctx, int64_t R, int M, int N, const float* G, int64_t ldg, const float* A, int64_t lda, float* C,
                     int64_t ldc, int accumulate);
