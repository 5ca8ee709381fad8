// sage_aggregate.cu - per-layer message-passing aggregate for sm_100a (fp32).
//
// Replaces torch_geometric SAGEConv / GCNConv as stacked by
// python/gigl/src/common/models/pyg/homogeneous.py:171-202 (GraphSAGE.init_conv_layers) and :527-542
// (TwoLayerGCN), called from graphsage_template_modeling_spec.py:305-311 / :565-577:
//   SAGE: out_i = Wl @ mean_{j->i} x_j + bl + Wr @ x_i          GCN: out_i = sum_j dinv_j dinv_i (W x_j) + b
//
// Round-1 structure: (1) gather-mean over CSR-by-dst rows, one warp per destination row, float4
// loads, sub-warp neighbour groups for narrow features, warp-shuffle segment reduce;
// (2) projection [agg | x] @ [Wl | Wr]^T + b as a shared-memory tiled fp32 GEMM with the bias /
// ReLU epilogue fused.  fp32 FFMA keeps the 1e-5 relative parity bound; the tensor-core
// (3xTF32 tcgen05) projection is the planned replacement.
#include <cuda_runtime.h>

#include "common.cuh"

namespace gigl {

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void add4(float4& a, const float4& b) {
    a.x += b.x;
    a.y += b.y;
    a.z += b.z;
    a.w += b.w;
}

// LPR lanes cover one 4*LPR-float chunk of a feature row; G = 32/LPR neighbours are read at once.
// MODE 0: mean (SAGE).  MODE 1: GCN - skip self loops, weight by dinv[src], add the implicit self
// loop, scale by dinv[row], add bias, optional relu (x is then x' = x W^T).
template <int LPR, int MODE>
__global__ void __launch_bounds__(256) gather_rows_kernel(int64_t n_rows, int F, const int64_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ col, const float* __restrict__ x,
                                                          float* __restrict__ out, const float* __restrict__ dinv,
                                                          const float* __restrict__ bias, int relu) {
    constexpr int G = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int sub = lane % LPR;
    const int g = lane / LPR;
    const int64_t beg = __ldg(rowptr + row);
    const int64_t end = __ldg(rowptr + row + 1);
    float scale;
    if (MODE == 0) {
        const int64_t d = end - beg;
        scale = 1.0f / (float)(d > 1 ? d : 1);
    } else {
        scale = __ldg(dinv + row);
    }
    for (int c0 = 0; c0 < F; c0 += LPR * 4) {
        const int c = c0 + sub * 4;
        const bool active = c < F;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t base = beg; base < end; base += 32) {
            const int cnt = (int)((end - base) < 32 ? (end - base) : 32);
            int32_t my = -1;
            float myw = 0.f;
            if (lane < cnt) {
                my = __ldg(col + base + lane);
                if (MODE == 1) myw = (my == (int32_t)row) ? 0.f : __ldg(dinv + my);
            }
            for (int t = 0; t < cnt; t += 4 * G) {
                float4 v[4];
                float w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = t + u * G + g;
                    const int32_t s = __shfl_sync(0xffffffffu, my, j & 31);
                    if (MODE == 1) w[u] = __shfl_sync(0xffffffffu, myw, j & 31);
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < cnt && active) v[u] = ldg_f4(x + (int64_t)s * F + c);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (MODE == 1) {
                        acc.x += w[u] * v[u].x;
                        acc.y += w[u] * v[u].y;
                        acc.z += w[u] * v[u].z;
                        acc.w += w[u] * v[u].w;
                    } else {
                        add4(acc, v[u]);
                    }
                }
            }
        }
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
        }
        if (g == 0 && active) {
            float4 r;
            if (MODE == 0) {
                r = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
            } else {
                const float4 self = ldg_f4(x + row * F + c);
                const float4 b = bias ? ldg_f4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                r.x = (acc.x + scale * self.x) * scale + b.x;
                r.y = (acc.y + scale * self.y) * scale + b.y;
                r.z = (acc.z + scale * self.z) * scale + b.z;
                r.w = (acc.w + scale * self.w) * scale + b.w;
                if (relu) {
                    r.x = fmaxf(r.x, 0.f);
                    r.y = fmaxf(r.y, 0.f);
                    r.z = fmaxf(r.z, 0.f);
                    r.w = fmaxf(r.w, 0.f);
                }
            }
            *reinterpret_cast<float4*>(out + row * F + c) = r;
        }
    }
}

// Scalar fallback for feature widths that are not a multiple of 4 (or unaligned bases).
template <int MODE>
__global__ void __launch_bounds__(256) gather_rows_scalar_kernel(int64_t n_rows, int F,
                                                                 const int64_t* __restrict__ rowptr,
                                                                 const int32_t* __restrict__ col,
                                                                 const float* __restrict__ x, float* __restrict__ out,
                                                                 const float* __restrict__ dinv,
                                                                 const float* __restrict__ bias, int relu) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t beg = __ldg(rowptr + row);
    const int64_t end = __ldg(rowptr + row + 1);
    float scale;
    if (MODE == 0) {
        const int64_t d = end - beg;
        scale = 1.0f / (float)(d > 1 ? d : 1);
    } else {
        scale = __ldg(dinv + row);
    }
    for (int c = lane; c < F; c += 32) {
        float acc = 0.f;
        for (int64_t e = beg; e < end; ++e) {
            const int32_t s = __ldg(col + e);
            if (MODE == 1) {
                if (s != (int32_t)row) acc += __ldg(dinv + s) * __ldg(x + (int64_t)s * F + c);
            } else {
                acc += __ldg(x + (int64_t)s * F + c);
            }
        }
        float r;
        if (MODE == 0) {
            r = acc * scale;
        } else {
            r = (acc + scale * __ldg(x + row * F + c)) * scale + (bias ? __ldg(bias + c) : 0.f);
            if (relu) r = fmaxf(r, 0.f);
        }
        out[row * F + c] = r;
    }
}

// deg_i = 1 + #non-loop in-edges; dinv = deg^-1/2   (gcn_norm with add_remaining_self_loops)
__global__ void gcn_dinv_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                float* __restrict__ dinv) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int64_t beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    int c = 0;
    for (int64_t e = beg + lane; e < end; e += 32) c += (__ldg(col + e) != (int32_t)row);
#pragma unroll
    for (int off = 16; off; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) dinv[row] = 1.0f / sqrtf((float)(c + 1));
}

// C[M,N] = [A0 | A1][M, K0+K1] @ [B0 | B1][N, K0+K1]^T + bias, optional relu.  fp32 FFMA.
// 64x64 block tile, BK = 16, 256 threads, 4x4 register tile per thread.
constexpr int BM = 64, BN = 64, BK = 16;
__global__ void __launch_bounds__(256) linear2_kernel(int64_t M, int N, int K0, int K1, const float* __restrict__ A0,
                                                      const float* __restrict__ A1, const float* __restrict__ B0,
                                                      const float* __restrict__ B1, const float* __restrict__ bias,
                                                      float* __restrict__ C, int relu) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int K = K0 + K1;
    // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread: row = tid/4, k = (tid%4)*4 .. +3
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            float a = 0.f, b = 0.f;
            const int64_t m = m0 + lr;
            if (m < M && k < K) a = (k < K0) ? __ldg(A0 + m * K0 + k) : __ldg(A1 + m * K1 + (k - K0));
            const int n = n0 + lr;
            if (n < N && k < K) b = (k < K0) ? __ldg(B0 + (int64_t)n * K0 + k) : __ldg(B1 + (int64_t)n * K1 + (k - K0));
            As[lk + q][lr] = a;
            Bs[lk + q][lr] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[m * N + n] = v;
        }
    }
}

// C[M, N] = A[M, K] @ W[N, K]^T + bias, optional relu; explicit leading dimensions, and M may live
// on the device (*m_dev, clamped to M) so a batch pipeline needs no host round trip.  fp32 FFMA.
__global__ void __launch_bounds__(256) linear_ld_kernel(const int32_t* __restrict__ m_dev, int64_t M, int N, int K,
                                                        const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ W, int64_t ldw,
                                                        const float* __restrict__ bias, float* __restrict__ C,
                                                        int64_t ldc, int relu) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    if (m_dev) {
        const int64_t md = *m_dev;
        if (md < M) M = md;
    }
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    if (m0 >= M) return;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.y * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            const int64_t m = m0 + lr;
            const int n = n0 + lr;
            As[lk + q][lr] = (m < M && k < K) ? __ldg(A + m * lda + k) : 0.f;
            Bs[lk + q][lr] = (n < N && k < K) ? __ldg(W + (int64_t)n * ldw + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            C[m * ldc + n] = v;
        }
    }
}

template <int MODE>
static int launch_gather(gigl_ctx* ctx, int64_t n_rows, int32_t F, const int64_t* rowptr, const int32_t* col,
                         const float* x, float* out, const float* dinv, const float* bias, int relu) {
    if (n_rows == 0 || F == 0) return GIGL_OK;
    const int wpb = 8;
    const int64_t blocks = ceil_div64(n_rows, wpb);
    if (blocks > 0x7fffffffLL) return gigl_fail(ctx, GIGL_E_INVALID, "too many rows for one launch");
    const bool vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                     (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    dim3 grid((unsigned)blocks), block(wpb * 32);
    if (!vec) {
        gather_rows_scalar_kernel<MODE><<<grid, block, 0, ctx->stream>>>(n_rows, F, rowptr, col, x, out, dinv, bias, relu);
    } else if (F <= 16) {
        gather_rows_kernel<4, MODE><<<grid, block, 0, ctx->stream>>>(n_rows, F, rowptr, col, x, out, dinv, bias, relu);
    } else if (F <= 32) {
        gather_rows_kernel<8, MODE><<<grid, block, 0, ctx->stream>>>(n_rows, F, rowptr, col, x, out, dinv, bias, relu);
    } else if (F <= 64) {
        gather_rows_kernel<16, MODE><<<grid, block, 0, ctx->stream>>>(n_rows, F, rowptr, col, x, out, dinv, bias, relu);
    } else {
        gather_rows_kernel<32, MODE><<<grid, block, 0, ctx->stream>>>(n_rows, F, rowptr, col, x, out, dinv, bias, relu);
    }
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

static int launch_linear2(gigl_ctx* ctx, int64_t M, int N, int K0, int K1, const float* A0, const float* A1,
                          const float* B0, const float* B1, const float* bias, float* C, int relu) {
    if (M == 0 || N == 0) return GIGL_OK;
    const int64_t gx = ceil_div64(M, BM);
    if (gx > 0x7fffffffLL) return gigl_fail(ctx, GIGL_E_INVALID, "too many rows for one launch");
    dim3 grid((unsigned)gx, (unsigned)((N + BN - 1) / BN));
    linear2_kernel<<<grid, 256, 0, ctx->stream>>>(M, N, K0, K1, A0, A1, B0, B1, bias, C, relu);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

}  // namespace gigl

int linear_dev_rows_launch(gigl_ctx* ctx, const int32_t* m_dev, int64_t m_cap, int N, int K, const float* A, int64_t lda,
                           const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int relu) {
    using namespace gigl;
    if (m_cap == 0 || N == 0) return GIGL_OK;
    const int64_t gx = ceil_div64(m_cap, BM);
    if (gx > 0x7fffffffLL) return gigl_fail(ctx, GIGL_E_INVALID, "too many rows for one launch");
    dim3 grid((unsigned)gx, (unsigned)((N + BN - 1) / BN));
    linear_ld_kernel<<<grid, 256, 0, ctx->stream>>>(m_dev, m_cap, N, K, A, lda, W, ldw, bias, C, ldc, relu);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

int gather_mean_launch(gigl_ctx* ctx, int64_t n_rows, int32_t F, const int64_t* rowptr, const int32_t* col,
                       const float* x, float* agg) {
    gigl_timed t(ctx, GIGL_T_GATHER_FULL);
    return gigl::launch_gather<0>(ctx, n_rows, F, rowptr, col, x, agg, nullptr, nullptr, 0);
}

int sage_conv_launch(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O, const int64_t* rowptr,
                     const int32_t* col, const float* x, const float* Wl, const float* bl, const float* Wr,
                     float* out, int32_t relu) {
    GIGL_CHECK(ctx, n >= 0 && n_rows_out >= 0 && n_rows_out <= n, "bad row counts");
    GIGL_CHECK(ctx, F >= 1 && O >= 1, "bad feature dims");
    if (n_rows_out == 0) return GIGL_OK;
    void* scratch = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_AGG, sizeof(float) * (size_t)n_rows_out * (size_t)F, &scratch);
    if (rc != GIGL_OK) return rc;
    float* agg = (float*)scratch;
    {
        gigl_timed t(ctx, GIGL_T_GATHER_FULL);
        rc = gigl::launch_gather<0>(ctx, n_rows_out, F, rowptr, col, x, agg, nullptr, nullptr, 0);
    }
    if (rc != GIGL_OK) return rc;
    gigl_timed t(ctx, GIGL_T_GEMM_FULL);
    return gigl::launch_linear2(ctx, n_rows_out, O, F, F, agg, x, Wl, Wr, bl, out, relu);
}

int gcn_conv_launch(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr, const int32_t* col,
                    const float* x, const float* W, const float* b, float* out, int32_t relu) {
    GIGL_CHECK(ctx, n >= 0 && F >= 1 && O >= 1, "bad sizes");
    if (n == 0) return GIGL_OK;
    // scratch = x' [n, O] followed by dinv [n] (x' start stays 16B aligned; dinv after a padded size)
    const size_t xp_elems = ((size_t)n * (size_t)O + 3) & ~(size_t)3;
    void* scratch = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_AGG, sizeof(float) * (xp_elems + (size_t)n), &scratch);
    if (rc != GIGL_OK) return rc;
    float* xp = (float*)scratch;
    float* dinv = xp + xp_elems;
    rc = gigl::launch_linear2(ctx, n, O, F, 0, x, nullptr, W, nullptr, nullptr, xp, 0);
    if (rc != GIGL_OK) return rc;
    const int wpb = 8;
    gigl::gcn_dinv_kernel<<<(unsigned)ceil_div64(n, wpb), wpb * 32, 0, ctx->stream>>>(n, rowptr, col, dinv);
    GIGL_LAUNCHED(ctx);
    return gigl::launch_gather<1>(ctx, n, O, rowptr, col, xp, out, dinv, b, relu);
}
