// gemm_tn_tcgen05.cu - the weight-gradient GEMM of the aggregate's backward pass on tcgen05 (sm_100a).
//
//   C[M, N] = sum_{r < R} G[r, M] * A[r, N]        (C = G^T A;  R = batch rows, the long dimension)
//
// This is d(lin_l.weight | lin_r.weight) = grad_out^T @ [mean | self] of torch_geometric SAGEConv as the
// reference trains it (loss.backward() in python/gigl/src/common/modeling_task_specs/
// graphsage_template_modeling_spec.py:299-367 and node_classification_modeling_task_spec.py:134-173).
//
// Both operands are stored row-major with the reduction index r as the SLOW dimension, i.e. they are "MN-major"
// for the tensor core: TMA boxes of [kTnBlockK rows x 32 floats] land in shared memory as 32-float chunks in the
// "128-byte swizzle, 32-byte atom" pattern (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) - the only shared-memory layout
// tcgen05 accepts for MN-major TF32 operands (plain SWIZZLE_128B silently yields zeros: measured, see
// scripts/tn_probe.cu) - and the UMMA descriptors are built with Major-MN / SWIZZLE_128B_BASE32B (LBO = chunk
// pitch, SBO = 512 bytes between 4-row atoms).  No transposed copy of the activations is ever made.
//
// Precision: 3xTF32 like the forward projection (hi*hi + lo*hi + hi*lo accumulated in fp32 in TMEM).
// Parallelism: split-K.  The reduction over R is cut into `splits` contiguous ranges, one CTA per
// (m tile, n tile, split); every CTA writes its fp32 partial tile to a workspace and a second kernel sums
// the partials in split order, so the result is run-to-run deterministic (no float atomics).
//
//   warp 0      : TMA producer        warp 1 : TMEM alloc + MMA issue (one thread)
//   warps 2..5  : epilogue (tcgen05.ld -> partial tile)
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace gigl {

constexpr int kTnBlockM = 128;
constexpr int kTnBlockK = 16;   // rows of R per stage
constexpr int kTnChunk = 32;    // floats per 128-byte swizzle span
constexpr int kTnThreads = 192;

struct TnParams {
    int64_t R;
    int M, N;
    int n_pad;       // UMMA N (multiple of 32, <= 256)
    int m_tiles, n_tiles, splits;
    int kb_per_split;
    int stages;
    uint32_t tmem_cols;
    float* partial;  // [splits][m_tiles * 128][n_tiles * n_pad]
};

__global__ void __launch_bounds__(kTnThreads, 1)
linear_tn_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
                        const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo, const TnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t chunk_bytes = kTnBlockK * 128;                 // one 32-float chunk x kTnBlockK rows
    constexpr uint32_t g_bytes = (kTnBlockM / kTnChunk) * chunk_bytes;  // 4 chunks
    const uint32_t a_chunks = (uint32_t)p.n_pad / kTnChunk;
    const uint32_t a_bytes = a_chunks * chunk_bytes;
    const uint32_t stage_bytes = 2 * g_bytes + 2 * a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * p.stages;
    const uint32_t bar_tfull = bar_empty + 8 * p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
    const uint32_t smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item
    int w = blockIdx.x;
    const int split = w % p.splits;
    w /= p.splits;
    const int nt = w % p.n_tiles;
    const int mt = w / p.n_tiles;
    const int64_t total_kb = (p.R + kTnBlockK - 1) / kTnBlockK;
    const int64_t kb0 = (int64_t)split * p.kb_per_split;
    int64_t kb1 = kb0 + p.kb_per_split;
    if (kb1 > total_kb) kb1 = total_kb;
    const int num_kb = (int)(kb1 > kb0 ? kb1 - kb0 : 0);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && num_kb > 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int m0 = mt * kTnBlockM, n0 = nt * p.n_pad;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                const uint32_t full = bar_full + 8 * stage;
                const uint32_t sg = smem_base + stage * stage_bytes;
                const int r0 = (int)((kb0 + kb) * kTnBlockK);
                mbar_expect_tx(full, stage_bytes);
#pragma unroll
                for (int c = 0; c < kTnBlockM / kTnChunk; ++c) {
                    tma_load_2d(sg + c * chunk_bytes, &tm_g_hi, full, m0 + c * kTnChunk, r0);
                    tma_load_2d(sg + g_bytes + c * chunk_bytes, &tm_g_lo, full, m0 + c * kTnChunk, r0);
                }
                const uint32_t sa = sg + 2 * g_bytes;
                for (uint32_t c = 0; c < a_chunks; ++c) {
                    tma_load_2d(sa + c * chunk_bytes, &tm_a_hi, full, n0 + (int)c * kTnChunk, r0);
                    tma_load_2d(sa + a_bytes + c * chunk_bytes, &tm_a_lo, full, n0 + (int)c * kTnChunk, r0);
                }
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && num_kb > 0) {
            // D = F32, A = B = TF32, both MN-major (bits 15 / 16), N = n_pad, M = 128
#ifndef TN_PROBE_MAJOR_BITS
#define TN_PROBE_MAJOR_BITS ((1u << 15) | (1u << 16))
#endif
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | TN_PROBE_MAJOR_BITS | ((uint32_t)(p.n_pad >> 3) << 17) |
                                   ((uint32_t)(kTnBlockM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                const uint32_t sg = smem_base + stage * stage_bytes;
                const uint32_t sa = sg + 2 * g_bytes;
#pragma unroll
                for (int k = 0; k < kTnBlockK / 8; ++k) {
                    const uint32_t off = (uint32_t)k * 1024;  // next 8 rows (two 4-row swizzle atoms) inside every chunk
                    const uint64_t d_gh = umma_desc_mn_sw128_32b(sg + off, chunk_bytes, 512);
                    const uint64_t d_gl = umma_desc_mn_sw128_32b(sg + g_bytes + off, chunk_bytes, 512);
                    const uint64_t d_ah = umma_desc_mn_sw128_32b(sa + off, chunk_bytes, 512);
                    const uint64_t d_al = umma_desc_mn_sw128_32b(sa + a_bytes + off, chunk_bytes, 512);
                    umma_tf32(tmem_base, d_gl, d_ah, idesc, (kb | k) != 0);
                    umma_tf32(tmem_base, d_gh, d_al, idesc, 1);
                    umma_tf32(tmem_base, d_gh, d_ah, idesc, 1);
                }
                umma_commit(bar_empty + 8 * stage);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            umma_commit(bar_tfull);
        }
    } else {
        const int quarter = warp & 3;
        const int64_t ldp = (int64_t)p.n_tiles * p.n_pad;
        float* prow = p.partial + ((int64_t)split * p.m_tiles * kTnBlockM + (int64_t)mt * kTnBlockM + quarter * 32 + lane) * ldp +
                      (int64_t)nt * p.n_pad;
        if (num_kb > 0) {
            mbar_wait(bar_tfull, 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
            for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
        } else {
            for (int c0 = 0; c0 < p.n_pad; c0 += 4) *reinterpret_cast<float4*>(prow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// C[m, n] (+)= sum_s partial[s][m][n] in split order; columns [0, n_split) go to C0, the rest to C1 (so one GEMM over
// [mean | self] fills lin_l.weight.grad and lin_r.weight.grad directly).
__global__ void tn_reduce_kernel(int M, int N, int splits, int64_t rows_pad, int64_t ldp, const float* __restrict__ partial,
                                 float* __restrict__ C0, int64_t ldc0, int n0_valid, int n_split, float* __restrict__ C1,
                                 int64_t ldc1, int n1_valid, int accumulate) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i - (int64_t)m * N);
        if (n < n_split ? n >= n0_valid : (n - n_split) >= n1_valid) continue;  // padding columns of the operands
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += partial[((int64_t)s * rows_pad + m) * ldp + n];
        float* dst = (n < n_split) ? (C0 + (int64_t)m * ldc0 + n) : (C1 + (int64_t)m * ldc1 + (n - n_split));
        *dst = accumulate ? (*dst + acc) : acc;
    }
}

// column sums of G[R, M] (the bias gradient): stage 1 = per-block partial sums over a row range, stage 2 = fixed-order sum
__global__ void colsum_partial_kernel(int64_t R, int M, const float* __restrict__ G, int64_t ldg, int64_t rows_per_block,
                                      float* __restrict__ partial) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    int64_t r1 = r0 + rows_per_block;
    if (r1 > R) r1 = R;
    for (int c = threadIdx.x; c < M; c += blockDim.x) {
        float acc = 0.f;
        for (int64_t r = r0; r < r1; ++r) acc += G[r * ldg + c];
        partial[(int64_t)blockIdx.x * M + c] = acc;
    }
}
__global__ void colsum_final_kernel(int M, int blocks, const float* __restrict__ partial, float* __restrict__ out, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    float acc = 0.f;
    for (int b = 0; b < blocks; ++b) acc += partial[(int64_t)b * M + c];
    out[c] = accumulate ? out[c] + acc : acc;
}

// 2-D fp32 tensor [rows, cols], pitch ld floats; box = [box_rows, 32 floats], 128-byte swizzle, OOB -> 0
static int make_map_tn(gigl_ctx* ctx, CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return gigl_fail(ctx, GIGL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTnChunk, (cuuint32_t)kTnBlockK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gigl_fail(ctx, GIGL_E_CUDA, "cuTensorMapEncodeTiled failed (pointer / pitch must be 16-byte aligned)");
    return GIGL_OK;
}

}  // namespace gigl

// G_hi / G_lo: [R, M] pitch ldg; A_hi / A_lo: [R, N] pitch lda (16-byte aligned, pitches % 4 == 0).
// C0 gets columns [0, n0_valid) (pitch ldc0), C1 the columns [n_split, n_split + n1_valid) (pitch ldc1); columns in between /
// beyond are operand padding and are dropped.  n_split = N and C1 = NULL for one output.
int linear_tn_tc_launch(gigl_ctx* ctx, int64_t R, int M, int N, const float* G_hi, const float* G_lo, int64_t ldg, const float* A_hi,
                        const float* A_lo, int64_t lda, float* C0, int64_t ldc0, int n0_valid, int n_split, float* C1, int64_t ldc1,
                        int n1_valid, int accumulate) {
    using namespace gigl;
    if (M == 0 || N == 0) return GIGL_OK;
    GIGL_CHECK(ctx, R >= 0 && R <= 0x7fffffffLL && ldg % 4 == 0 && lda % 4 == 0, "weight-gradient GEMM needs pitches that are multiples of 4 floats");
    GIGL_CHECK(ctx, n_split >= 0 && n_split <= N && (n_split == N || C1 != nullptr), "bad output split");
    TnParams p{};
    p.R = R;
    p.M = M;
    p.N = N;
    p.m_tiles = (M + kTnBlockM - 1) / kTnBlockM;
    p.n_tiles = (N + 255) / 256;
    const int n_per = (N + p.n_tiles - 1) / p.n_tiles;
    p.n_pad = (n_per + 31) & ~31;
    const int64_t total_kb = R > 0 ? (R + kTnBlockK - 1) / kTnBlockK : 0;
    int splits = ctx->sm_count / (p.m_tiles * p.n_tiles);
    if (splits < 1) splits = 1;
    if ((int64_t)splits > total_kb) splits = (int)(total_kb > 0 ? total_kb : 1);
    p.kb_per_split = (int)((total_kb + splits - 1) / splits);
    if (p.kb_per_split < 1) p.kb_per_split = 1;
    splits = (int)((total_kb + p.kb_per_split - 1) / p.kb_per_split);
    if (splits < 1) splits = 1;
    p.splits = splits;
    const size_t stage_bytes = 2 * (size_t)(kTnBlockM + p.n_pad) * kTnBlockK * 4;
    int stages = (int)((220 * 1024 - 1024 - 256) / stage_bytes);
    if (stages > 6) stages = 6;
    GIGL_CHECK(ctx, stages >= 2, "tile does not fit in shared memory");
    p.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)p.n_pad) cols <<= 1;
    p.tmem_cols = cols;
    const int64_t rows_pad = (int64_t)p.m_tiles * kTnBlockM, ldp = (int64_t)p.n_tiles * p.n_pad;
    void* ws = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_WORK, sizeof(float) * (size_t)splits * rows_pad * ldp, &ws);
    if (rc != GIGL_OK) return rc;
    p.partial = (float*)ws;
    CUtensorMap tg_hi, tg_lo, ta_hi, ta_lo;
    const int64_t Rm = R > 0 ? R : 1;
    if ((rc = make_map_tn(ctx, &tg_hi, G_hi, Rm, M, ldg)) != GIGL_OK) return rc;
    if ((rc = make_map_tn(ctx, &tg_lo, G_lo, Rm, M, ldg)) != GIGL_OK) return rc;
    if ((rc = make_map_tn(ctx, &ta_hi, A_hi, Rm, N, lda)) != GIGL_OK) return rc;
    if ((rc = make_map_tn(ctx, &ta_lo, A_lo, Rm, N, lda)) != GIGL_OK) return rc;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        GIGL_CUDA(ctx, cudaFuncSetAttribute(linear_tn_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = p.m_tiles * p.n_tiles * splits;
    linear_tn_tf32x3_kernel<<<grid, kTnThreads, smem, ctx->stream>>>(tg_hi, tg_lo, ta_hi, ta_lo, p);
    GIGL_LAUNCHED(ctx);
    const int64_t total = (int64_t)M * N;
    int64_t rb = ceil_div64(total, 256);
    if (rb > ctx->sm_count * 8) rb = ctx->sm_count * 8;
    tn_reduce_kernel<<<(unsigned)rb, 256, 0, ctx->stream>>>(M, N, splits, rows_pad, ldp, p.partial, C0, ldc0, n0_valid, n_split, C1, ldc1, n1_valid,
                                                                 accumulate);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

// out[c] (+)= sum_r G[r, c]; deterministic two-stage reduction.  Uses GIGL_SLOT_SORT as workspace.
int colsum_launch(gigl_ctx* ctx, int64_t R, int M, const float* G, int64_t ldg, float* out, int accumulate) {
    using namespace gigl;
    if (M == 0) return GIGL_OK;
    int blocks = ctx->sm_count * 4;
    if ((int64_t)blocks > R) blocks = (int)(R > 0 ? R : 1);
    const int64_t rpb = R > 0 ? ceil_div64(R, blocks) : 1;
    blocks = (int)(R > 0 ? ceil_div64(R, rpb) : 1);
    void* ws = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_SORT, sizeof(float) * (size_t)blocks * M, &ws);
    if (rc != GIGL_OK) return rc;
    colsum_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(R, M, G, ldg, rpb, (float*)ws);
    GIGL_LAUNCHED(ctx);
    colsum_final_kernel<<<(M + 127) / 128, 128, 0, ctx->stream>>>(M, blocks, (const float*)ws, out, accumulate);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}
