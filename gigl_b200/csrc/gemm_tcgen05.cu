// gemm_tcgen05.cu - the layer's weight projection on the 5th-gen tensor cores (sm_100a).
//
//   C[M, N] = A[M, K] @ W[N, K]^T + bias  (optional ReLU),  fp32 in / fp32 out,
//
// i.e. F.linear of torch_geometric SAGEConv's lin_l / lin_r (fused as [mean | self] @ [Wl | Wr]^T;
// reference call site python/gigl/src/common/models/pyg/homogeneous.py:192-201) and GCNConv's lin.
//
// Precision: the parity bound is 1e-5 relative on fp32 embeddings, which single-pass TF32
// (10-bit mantissa) cannot meet, so every operand is split into two TF32 terms
// (x ~ hi + lo, hi = x rounded to TF32, lo = (x - hi) rounded to TF32; common.cuh gigl_split_tf32) and
// three tensor-core products are accumulated in fp32 in TMEM:  hi*hi + lo*hi + hi*lo  (3xTF32,
// relative error ~2^-21).  W is split once at model-creation time.  A arrives either already split (two tensors, the
// full-graph forms) or RAW (split_a: the batch path) - then the fp32 tile TMA drops into shared memory is split in
// place by four converter warps (hi over the raw tile, lo beside it) before the MMA warp sees it, so the intermediate
// [mean | self] rows cross HBM once in each direction instead of twice:
//
//   warp 0      : TMA producer - 128-byte-swizzled K-major tiles per stage (A_hi, A_lo | A raw; W_hi, W_lo)
//   warp 1      : TMEM allocation + single-thread tcgen05.mma issue (kind::tf32, M=128, N<=256, K=8)
//   warps 2..5  : epilogue - tcgen05.ld the accumulator (each warp its own 32-lane quarter),
//                 bias + ReLU, fp32 rows to global memory
//   warps 6..9  : (split_a only) TF32 split of the A tile, elementwise, so the swizzle is untouched;
//                 fence.proxy.async hands the tile from the generic to the async proxy
//
// Persistent over the M tiles (grid = min(tiles, SMs)), accumulators double-buffered in TMEM so
// the epilogue of tile t overlaps the MMAs of tile t+1, smem ring of 2..4 stages.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace gigl {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;              // 32 fp32 = 128 bytes = one swizzle span
constexpr int kUmmaK = 8;                // tf32
constexpr int kGemmThreads = 192;
constexpr int kGemmThreadsSplit = 320;   // + 4 converter warps
constexpr int kEpiPitch = 36;             // floats per row of an epilogue warp's 32 x 32 staging tile
constexpr size_t kEpiBytes = 4 * 32 * kEpiPitch * sizeof(float);


struct GemmParams {
    int64_t M;
    int N, K;
    int n_pad;        // UMMA N of one tile (multiple of 16, <= 256)
    int n_tiles_n;    // tiles along N
    int stages;
    int nbuf;         // TMEM accumulator buffers (2 = the epilogue of tile t overlaps the MMAs of tile t + 1)
    int split_acc;    // 1 = the two small products (lo*hi, hi*lo) accumulate in their own TMEM columns (long K)
    uint32_t tmem_cols;
    const float* bias;
    float* C;
    int64_t ldc;
    int relu;
    int split_a;           // A is raw fp32 (tm_a_hi maps it, tm_a_lo is unused): split in shared memory
    const int32_t* m_dev;  // optional device-side row count (<= M): the host sized buffers and maps by an upper bound
    int dbg;  // GIGL_GEMM_DBG (timing experiments only): 1 = no MMA, 2 = no W loads, 4 = no A loads, 8 = no epilogue stores
};

__global__ void __launch_bounds__(kGemmThreadsSplit, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                     const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                     const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment of every tile is required by the 128-byte swizzle
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = kBlockM * kBlockK * 4;          // 16 KB
    const uint32_t w_bytes = (uint32_t)p.n_pad * kBlockK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * w_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    // bars: full[stages] (A tiles) | empty[stages] | tmem_full[2] | tmem_empty[2] | conv[stages] | fullw[stages] (W tiles)
    // A and W land on separate barriers: the converter warps split A while the (larger) W tiles are still in flight
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * p.stages;
    const uint32_t bar_tfull = bar_empty + 8 * p.stages;
    const uint32_t bar_tempty = bar_tfull + 16;
    const uint32_t bar_conv = bar_tempty + 16;
    const uint32_t bar_fullw = bar_conv + 8 * p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * p.stages + 4);
    const uint32_t smem_base = smem_u32(smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int64_t M = p.M;
    if (p.m_dev != nullptr) {
        const int64_t m = *p.m_dev;
        if (m < M) M = m;
    }
    const int64_t n_tiles_m = (M + kBlockM - 1) / kBlockM;
    const int64_t n_tiles = n_tiles_m * p.n_tiles_n;
    const int num_kb = (p.K + kBlockK - 1) / kBlockK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 4);  // one arrive per converter warp
            mbar_init(bar_fullw + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, 4);  // one arrive per epilogue warp
        }
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
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (int)((tile / p.n_tiles_n) * kBlockM);
                const int n0 = (int)((tile % p.n_tiles_n) * p.n_pad);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t full = bar_full + 8 * stage, fullw = bar_fullw + 8 * stage;
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    mbar_expect_tx(full, (p.dbg & 4) ? 0 : (p.split_a ? a_bytes : 2 * a_bytes));
                    if (!(p.dbg & 4)) {
                        tma_load_2d(sa, &tm_a_hi, full, kb * kBlockK, m0);
                        if (!p.split_a) tma_load_2d(sa + a_bytes, &tm_a_lo, full, kb * kBlockK, m0);
                    }
                    mbar_expect_tx(fullw, (p.dbg & 2) ? 0 : 2 * w_bytes);
                    if (!(p.dbg & 2)) {
                        tma_load_2d(sa + 2 * a_bytes, &tm_w_hi, fullw, kb * kBlockK, n0);
                        tma_load_2d(sa + 2 * a_bytes + w_bytes, &tm_w_lo, fullw, kb * kBlockK, n0);
                    }
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = n_pad, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = p.nbuf == 2 ? (it & 1) : 0;
                const uint32_t use = p.nbuf == 2 ? (it >> 1) : it;
                mbar_wait(bar_tempty + 8 * acc, (use & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc_cols = (uint32_t)p.n_pad << p.split_acc;
                const uint32_t tmem_d = tmem_base + acc * acc_cols;
                const uint32_t tmem_s = tmem_d + (uint32_t)p.n_pad;  // small-term accumulator (split_acc only)
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait((p.split_a ? bar_conv : bar_full) + 8 * stage, phase);
                    mbar_wait(bar_fullw + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint64_t d_ah = umma_desc_sw128(sa), d_al = umma_desc_sw128(sa + a_bytes);
                    const uint64_t d_wh = umma_desc_sw128(sa + 2 * a_bytes), d_wl = umma_desc_sw128(sa + 2 * a_bytes + w_bytes);
#pragma unroll
                    for (int k = 0; k < ((p.dbg & 1) ? 0 : kBlockK / kUmmaK); ++k) {
                        const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);  // +32 bytes inside the swizzle span
                        if (p.split_acc) {
                            // the accumulator add truncates: keeping the 2^-11-sized terms out of the big sum costs it
                            // one rounding per k step instead of three
                            umma_tf32(tmem_s, d_al + adv, d_wh + adv, idesc, (kb | k) != 0);
                            umma_tf32(tmem_s, d_ah + adv, d_wl + adv, idesc, 1);
                            umma_tf32(tmem_d, d_ah + adv, d_wh + adv, idesc, (kb | k) != 0);
                        } else {
                            umma_tf32(tmem_d, d_al + adv, d_wh + adv, idesc, (kb | k) != 0);
                            umma_tf32(tmem_d, d_ah + adv, d_wl + adv, idesc, 1);
                            umma_tf32(tmem_d, d_ah + adv, d_wh + adv, idesc, 1);
                        }
                    }
                    umma_commit(bar_empty + 8 * stage);  // frees the smem stage when these MMAs retire
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 6) {
        // ===== converter warps 6..9 (split_a): raw fp32 A tile -> TF32 hi (in place) + lo =====
        if (p.split_a) {
            const int t = threadIdx.x - 192;  // 0..127: eight float4 each, consecutive threads on consecutive 16 bytes
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    float4* hi = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes);
                    float4* lo = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + a_bytes);
#pragma unroll
                    for (int q = 0; q < (int)(kBlockM * kBlockK / 4 / 128); ++q) {
                        const int idx = q * 128 + t;
                        const float4 v = hi[idx];
                        float4 h, l;
                        gigl_split_tf32(v.x, h.x, l.x);
                        gigl_split_tf32(v.y, h.y, l.y);
                        gigl_split_tf32(v.z, h.z, l.z);
                        gigl_split_tf32(v.w, h.w, l.w);
                        hi[idx] = h;
                        lo[idx] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_conv + 8 * stage);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        // tcgen05.ld hands every lane one accumulator ROW, so storing straight from the registers writes 32 different
        // rows per instruction, 16 bytes each (measured: the stores alone were 0.23 ms of a 0.38 ms projection).  Each
        // warp therefore turns its 32 x 32 chunk around through a private shared-memory tile (row pitch 36 floats: the
        // float4 accesses of both directions are conflict-free) and writes 4 rows x 128 contiguous bytes per instruction.
        const int quarter = warp & 3;
        const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        float* stg = reinterpret_cast<float*>(smem + (size_t)p.stages * stage_bytes + 256) + (warp - 2) * (32 * kEpiPitch);
        const int rr = lane >> 3, cc = (lane & 7) * 4;  // read-back role: row within a group of 4, first of 4 columns
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t acc = p.nbuf == 2 ? (it & 1) : 0;
            const uint32_t use = p.nbuf == 2 ? (it >> 1) : it;
            const int64_t m0 = (tile / p.n_tiles_n) * kBlockM;
            const int n0 = (int)((tile % p.n_tiles_n) * p.n_pad);
            mbar_wait(bar_tfull + 8 * acc, use & 1);
            tc_fence_after();
            const int64_t row0 = m0 + quarter * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * ((uint32_t)p.n_pad << p.split_acc);
            for (int c0 = 0; c0 < p.n_pad; c0 += 32) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int cb = c0 + half * 16;
                    if (cb >= p.n_pad) break;  // n_pad is a multiple of 16, not of 32
                    uint32_t v[16];
                    tmem_ld16(taddr + cb, v);
                    if (p.split_acc) {
                        uint32_t sm[16];
                        tmem_ld16(taddr + p.n_pad + cb, sm);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(sm[j]));
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<uint4*>(stg + lane * kEpiPitch + half * 16 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
                __syncwarp();
                const int col = n0 + c0 + cc;     // first of this lane's 4 output columns
                const int ncol = p.N - col;       // valid columns from there
                if (!(p.dbg & 8)) {
                    if (vec_ok && ncol >= 4 && c0 + cc < p.n_pad) {
                        const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int r = 0; r < 32; r += 4) {
                            const int64_t row = row0 + r + rr;
                            float4 o = *reinterpret_cast<const float4*>(stg + (r + rr) * kEpiPitch + cc);
                            o.x += bv.x;
                            o.y += bv.y;
                            o.z += bv.z;
                            o.w += bv.w;
                            if (p.relu) {
                                o.x = fmaxf(o.x, 0.f);
                                o.y = fmaxf(o.y, 0.f);
                                o.z = fmaxf(o.z, 0.f);
                                o.w = fmaxf(o.w, 0.f);
                            }
                            if (row < M) *reinterpret_cast<float4*>(p.C + row * p.ldc + col) = o;
                        }
                    } else if (ncol > 0 && c0 + cc < p.n_pad) {
                        for (int r = 0; r < 32; r += 4) {
                            const int64_t row = row0 + r + rr;
                            if (row >= M) continue;
                            for (int j = 0; j < 4 && j < ncol; ++j) {
                                float f = stg[(r + rr) * kEpiPitch + cc + j] + (p.bias ? __ldg(p.bias + col + j) : 0.f);
                                if (p.relu) f = fmaxf(f, 0.f);
                                p.C[row * p.ldc + col + j] = f;
                            }
                        }
                    }
                }
                __syncwarp();  // the tile is rewritten by the next chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// x -> (hi, lo): hi = x rounded to the nearest TF32 value, lo = (x - hi) rounded to TF32.
__global__ void split_tf32_kernel(int64_t rows, int cols, const float* __restrict__ x, int64_t ldx,
                                  float* __restrict__ hi, float* __restrict__ lo, int64_t ldo) {
    const int64_t total = rows * cols;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        const float v = x[r * ldx + c];
        gigl_split_tf32(v, hi[r * ldo + c], lo[r * ldo + c]);
    }
}


// 2-D fp32 tensor [rows, cols] with row pitch ld (floats): box = [box_rows, 32 floats], 128-byte swizzle, OOB -> 0
static int make_map(gigl_ctx* ctx, CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    bool reused) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return gigl_fail(ctx, GIGL_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, reused ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return gigl_fail(ctx, GIGL_E_CUDA, "cuTensorMapEncodeTiled failed (pointer / pitch must be 16-byte aligned)");
    return GIGL_OK;
}

}  // namespace gigl

// A_hi / A_lo: [M, K] with pitch lda; W_hi / W_lo: [N, K] with pitch ldw; all 16-byte aligned, pitches % 4 == 0.
int linear_tc_launch(gigl_ctx* ctx, int64_t M, int N, int K, const float* A_hi, const float* A_lo, int64_t lda,
                     const float* W_hi, const float* W_lo, int64_t ldw, const float* bias, float* C, int64_t ldc, int relu) {
    return linear_tc_launch_ex(ctx, M, nullptr, N, K, A_hi, A_lo, lda, W_hi, W_lo, ldw, bias, C, ldc, relu);
}

// A_lo == nullptr: A_hi is the RAW fp32 operand, split inside the kernel.  m_dev: optional device-side row count <= M
// (M then is the bound the buffers were sized by; rows at and beyond *m_dev are neither computed nor stored).
int linear_tc_launch_ex(gigl_ctx* ctx, int64_t M, const int32_t* m_dev, int N, int K, const float* A_hi, const float* A_lo, int64_t lda,
                        const float* W_hi, const float* W_lo, int64_t ldw, const float* bias, float* C, int64_t ldc, int relu) {
    using namespace gigl;
    if (M == 0 || N == 0) return GIGL_OK;
    GIGL_CHECK(ctx, K >= 1 && lda % 4 == 0 && ldw % 4 == 0, "tensor-core projection needs pitches that are multiples of 4 floats");
    GIGL_CHECK(ctx, M <= 0x7fffffffLL, "too many rows");
    GemmParams p{};
    p.M = M;
    p.N = N;
    p.K = K;
    const int n_cap = 256;
    p.n_tiles_n = (N + n_cap - 1) / n_cap;
    const int n_per = (N + p.n_tiles_n - 1) / p.n_tiles_n;
    p.n_pad = (n_per + 15) & ~15;
    const size_t stage_bytes = 2 * (size_t)kBlockM * kBlockK * 4 + 2 * (size_t)p.n_pad * kBlockK * 4;
    int stages = (int)((227 * 1024 - 1024 - 256 - kEpiBytes) / stage_bytes);
    if (stages > 4) stages = 4;
    GIGL_CHECK(ctx, stages >= 2, "tile does not fit in shared memory");
    p.stages = stages;
    // long reductions: separate accumulator for the small products (see the MMA issue loop); TMEM has 512 columns
    // ... and for narrow outputs, where the second accumulator still leaves room for double buffering (2 x 2 x 128 columns):
    // the tensor core's fp32 accumulate truncates, a one-sided error that grows with the number of accumulation steps
    // (measured: 1.1e-5 absolute on a 0.45-sized output of a K = 512 projection, nine times the fp32 oracle's worst element)
    p.split_acc = (K > 512 || p.n_pad <= 128) ? 1 : 0;
    p.nbuf = (2 * (p.n_pad << p.split_acc) <= 512) ? 2 : 1;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.nbuf * (p.n_pad << p.split_acc))) cols <<= 1;
    p.tmem_cols = cols;
    p.bias = bias;
    p.C = C;
    p.ldc = ldc;
    p.relu = relu;
    p.split_a = A_lo == nullptr ? 1 : 0;
    p.m_dev = m_dev;
    static const int dbg = getenv("GIGL_GEMM_DBG") ? atoi(getenv("GIGL_GEMM_DBG")) : 0;
    p.dbg = dbg;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    int rc;
    if ((rc = make_map(ctx, &ta_hi, A_hi, M, K, lda, kBlockM, false)) != GIGL_OK) return rc;
    if ((rc = make_map(ctx, &ta_lo, A_lo ? A_lo : A_hi, M, K, lda, kBlockM, false)) != GIGL_OK) return rc;
    if ((rc = make_map(ctx, &tw_hi, W_hi, N, K, ldw, p.n_pad, true)) != GIGL_OK) return rc;
    if ((rc = make_map(ctx, &tw_lo, W_lo, N, K, ldw, p.n_pad, true)) != GIGL_OK) return rc;
    const size_t smem = (size_t)stages * stage_bytes + 1024 /*alignment slack*/ + 256 /*barriers + tmem slot*/ + kEpiBytes;
    static bool attr_set = false;
    if (!attr_set) {
        GIGL_CUDA(ctx, cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int64_t n_tiles = ((M + kBlockM - 1) / kBlockM) * p.n_tiles_n;
    const int grid = (int)(n_tiles < ctx->sm_count ? n_tiles : ctx->sm_count);
    linear_tf32x3_kernel<<<grid, p.split_a ? kGemmThreadsSplit : kGemmThreads, smem, ctx->stream>>>(ta_hi, ta_lo, tw_hi, tw_lo, p);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

int split_tf32_launch(gigl_ctx* ctx, int64_t rows, int cols, const float* x, int64_t ldx, float* hi, float* lo, int64_t ldo) {
    if (rows == 0 || cols == 0) return GIGL_OK;
    int64_t g = ceil_div64(rows * cols, 256);
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    gigl::split_tf32_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, ctx->stream>>>(rows, cols, x, ldx, hi, lo, ldo);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}
