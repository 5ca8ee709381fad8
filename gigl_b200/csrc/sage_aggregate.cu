// sage_aggregate.cu - per-layer message-passing aggregate for sm_100a (fp32).
//
// Replaces torch_geometric SAGEConv / GCNConv as stacked by
// python/gigl/src/common/models/pyg/homogeneous.py:171-202 (GraphSAGE.init_conv_layers) and :527-542
// (TwoLayerGCN), called from graphsage_template_modeling_spec.py:305-311 / :565-577:
//   SAGE: out_i = Wl @ mean_{j->i} x_j + bl + Wr @ x_i          GCN: out_i = sum_j dinv_j dinv_i (W x_j) + b
//
// Structure: (1) gather-mean over CSR-by-dst rows, one warp per destination row, float4 loads, sub-warp neighbour
// groups for narrow features, warp-shuffle segment reduce, written as [mean | self]; (2) projection
// [mean | self] @ [Wl | Wr]^T + b on tcgen05 as 3xTF32 (gemm_tcgen05.cu) with the bias / ReLU epilogue fused;
// (3) backward: transposed gather for the input gradient, tcgen05 TN GEMM for the weight gradients
// (gemm_tn_tcgen05.cu).
#include <cuda_runtime.h>

#include <initializer_list>

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
// MODE 2: transposed aggregate of the backward pass - sum_j w[j] * x[j] over the row, where only sources j < w_rows
// count (w = dinv, the 1/deg of the forward rows), plus `bias` used as a per-row addend [w_rows, F] with pitch ldb
// (the self-path gradient); rows >= w_rows get no addend.
// x rows have pitch ldx, out rows pitch ldo (floats).
//
// Power-law graphs: a row with 10^5 edges would serialise on one warp (measured 1.9 TB/s on the products-like graph
// with a 91 701-edge hub), so rows longer than kRowSplit are pushed to a work list and summed by gather_heavy_kernel,
// one CTA per row: every warp sums a fixed contiguous slice and the slices are added in slice order, so the result
// does not depend on scheduling.
constexpr int kRowSplit = 1024;
constexpr int kHeavyWarpsAgg = 16;
constexpr int kGatherDepth = 8;  // independent 16-byte loads in flight per lane

struct GatherArgs {
    int64_t n_rows;
    int F;
    const int64_t* rowptr;
    const int32_t* col;
    const float* x;
    int64_t ldx;
    float* out;
    int64_t ldo;
    const float* dinv;
    const float* bias;
    int relu;
    int64_t w_rows;
    int64_t ldb;
    int32_t* heavy_ctr;   // [0] = rows deferred
    int32_t* heavy_rows;  // [heavy_cap]
    int32_t heavy_cap;
};

// Sum over edges [beg, end) of the row for feature columns [c, c + 4) of this lane (lane group g takes every G-th source).
template <int LPR, int MODE>
__device__ __forceinline__ float4 gather_span(const GatherArgs& a, int64_t row, int64_t beg, int64_t end, int c, bool active,
                                              int lane, int g) {
    constexpr int G = 32 / LPR;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t base = beg; base < end; base += 32) {
        const int cnt = (int)((end - base) < 32 ? (end - base) : 32);
        int32_t my = -1;
        float myw = 0.f;
        if (lane < cnt) {
            my = __ldg(a.col + base + lane);
            if (MODE == 1) myw = (my == (int32_t)row) ? 0.f : __ldg(a.dinv + my);
            if (MODE == 2) {
                if (my < a.w_rows) {
                    myw = __ldg(a.dinv + my);
                } else {
                    my = -1;
                }
            }
        }
        for (int t = 0; t < cnt; t += kGatherDepth * G) {
            float4 v[kGatherDepth];
            float w[kGatherDepth];
#pragma unroll
            for (int u = 0; u < kGatherDepth; ++u) {
                const int j = t + u * G + g;
                const int32_t s = __shfl_sync(0xffffffffu, my, j & 31);
                if (MODE != 0) w[u] = __shfl_sync(0xffffffffu, myw, j & 31);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < cnt && active && s >= 0) v[u] = ldg_f4(a.x + (int64_t)s * a.ldx + c);
            }
#pragma unroll
            for (int u = 0; u < kGatherDepth; ++u) {
                if (MODE != 0) {
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
    return acc;
}

// Row epilogue on one 4-column slice (acc = the reduced neighbour sum).
template <int MODE>
__device__ __forceinline__ float4 finish_row(const GatherArgs& a, int64_t row, int c, float4 acc, float scale) {
    float4 r;
    if (MODE == 0) {
        r = make_float4(acc.x * scale, acc.y * scale, acc.z * scale, acc.w * scale);
    } else if (MODE == 2) {
        r = acc;
        if (a.bias != nullptr && row < a.w_rows) add4(r, ldg_f4(a.bias + row * a.ldb + c));
    } else {
        const float4 self = ldg_f4(a.x + row * a.ldx + c);
        const float4 b = a.bias ? ldg_f4(a.bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        r.x = (acc.x + scale * self.x) * scale + b.x;
        r.y = (acc.y + scale * self.y) * scale + b.y;
        r.z = (acc.z + scale * self.z) * scale + b.z;
        r.w = (acc.w + scale * self.w) * scale + b.w;
        if (a.relu) {
            r.x = fmaxf(r.x, 0.f);
            r.y = fmaxf(r.y, 0.f);
            r.z = fmaxf(r.z, 0.f);
            r.w = fmaxf(r.w, 0.f);
        }
    }
    return r;
}

template <int MODE>
__device__ __forceinline__ float row_scale(const GatherArgs& a, int64_t row, int64_t beg, int64_t end) {
    if (MODE == 0) {
        const int64_t d = end - beg;
        return 1.0f / (float)(d > 1 ? d : 1);
    }
    if (MODE == 1) return __ldg(a.dinv + row);
    return 1.0f;
}

template <int LPR, int MODE>
__global__ void __launch_bounds__(256) gather_rows_kernel(const GatherArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.n_rows) return;
    const int sub = lane % LPR;
    const int g = lane / LPR;
    const int64_t beg = __ldg(a.rowptr + row);
    const int64_t end = __ldg(a.rowptr + row + 1);
    if (end - beg > kRowSplit && a.heavy_rows != nullptr) {
        int slot = 0;
        if (lane == 0) slot = atomicAdd(a.heavy_ctr, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot < a.heavy_cap) {
            if (lane == 0) a.heavy_rows[slot] = (int32_t)row;
            return;
        }
        // list full: this warp does the long row alone (still exact)
    }
    const float scale = row_scale<MODE>(a, row, beg, end);
    for (int c0 = 0; c0 < a.F; c0 += LPR * 4) {
        const int c = c0 + sub * 4;
        const bool active = c < a.F;
        float4 acc = gather_span<LPR, MODE>(a, row, beg, end, c, active, lane, g);
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
        }
        if (g == 0 && active) *reinterpret_cast<float4*>(a.out + row * a.ldo + c) = finish_row<MODE>(a, row, c, acc, scale);
    }
}

// One CTA per deferred row; warp w sums edge slice w of kHeavyWarpsAgg (128 feature columns per pass), the slices are
// added in slice order through shared memory.
template <int MODE>
__global__ void __launch_bounds__(kHeavyWarpsAgg * 32) gather_heavy_kernel(const GatherArgs a) {
    __shared__ float4 s_part[kHeavyWarpsAgg][32];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    int n_heavy = *a.heavy_ctr;
    if (n_heavy > a.heavy_cap) n_heavy = a.heavy_cap;
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int64_t row = a.heavy_rows[h];
        const int64_t beg = __ldg(a.rowptr + row);
        const int64_t end = __ldg(a.rowptr + row + 1);
        const int64_t per = (((end - beg) + kHeavyWarpsAgg - 1) / kHeavyWarpsAgg + 31) & ~(int64_t)31;
        const int64_t b0 = min(end, beg + per * w), b1 = min(end, b0 + per);
        const float scale = row_scale<MODE>(a, row, beg, end);
        for (int c0 = 0; c0 < a.F; c0 += 128) {
            const int c = c0 + lane * 4;
            const bool active = c < a.F;
            s_part[w][lane] = gather_span<32, MODE>(a, row, b0, b1, c, active, lane, 0);
            __syncthreads();
            if (w == 0 && active) {
                float4 acc = s_part[0][lane];
#pragma unroll
                for (int k = 1; k < kHeavyWarpsAgg; ++k) add4(acc, s_part[k][lane]);
                *reinterpret_cast<float4*>(a.out + row * a.ldo + c) = finish_row<MODE>(a, row, c, acc, scale);
            }
            __syncthreads();
        }
    }
}

// Scalar fallback for feature widths that are not a multiple of 4 (or unaligned bases).
template <int MODE>
__global__ void __launch_bounds__(256) gather_rows_scalar_kernel(int64_t n_rows, int F,
                                                                 const int64_t* __restrict__ rowptr,
                                                                 const int32_t* __restrict__ col,
                                                                 const float* __restrict__ x, int64_t ldx,
                                                                 float* __restrict__ out, int64_t ldo,
                                                                 const float* __restrict__ dinv,
                                                                 const float* __restrict__ bias, int relu, int64_t w_rows,
                                                                 int64_t ldb) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int64_t beg = __ldg(rowptr + row);
    const int64_t end = __ldg(rowptr + row + 1);
    float scale;
    if (MODE == 0) {
        const int64_t d = end - beg;
        scale = 1.0f / (float)(d > 1 ? d : 1);
    } else if (MODE == 1) {
        scale = __ldg(dinv + row);
    } else {
        scale = 1.0f;
    }
    for (int c = lane; c < F; c += 32) {
        float acc = 0.f;
        for (int64_t e = beg; e < end; ++e) {
            const int32_t s = __ldg(col + e);
            if (MODE == 1) {
                if (s != (int32_t)row) acc += __ldg(dinv + s) * __ldg(x + (int64_t)s * ldx + c);
            } else if (MODE == 2) {
                if (s < w_rows) acc += __ldg(dinv + s) * __ldg(x + (int64_t)s * ldx + c);
            } else {
                acc += __ldg(x + (int64_t)s * ldx + c);
            }
        }
        float r;
        if (MODE == 0) {
            r = acc * scale;
        } else if (MODE == 2) {
            r = acc + ((bias != nullptr && row < w_rows) ? __ldg(bias + row * ldb + c) : 0.f);
        } else {
            r = (acc + scale * __ldg(x + row * ldx + c)) * scale + (bias ? __ldg(bias + c) : 0.f);
            if (relu) r = fmaxf(r, 0.f);
        }
        out[row * ldo + c] = r;
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

template <int MODE>
static int launch_gather(gigl_ctx* ctx, int64_t n_rows, int32_t F, const int64_t* rowptr, const int32_t* col,
                         const float* x, float* out, const float* dinv, const float* bias, int relu, int64_t ldx = 0,
                         int64_t ldo = 0, int64_t w_rows = 0, int64_t ldb = 0) {
    if (n_rows == 0 || F == 0) return GIGL_OK;
    if (ldx == 0) ldx = F;
    if (ldo == 0) ldo = F;
    if (ldb == 0) ldb = F;
    const int wpb = 8;
    const int64_t blocks = ceil_div64(n_rows, wpb);
    if (blocks > 0x7fffffffLL) return gigl_fail(ctx, GIGL_E_INVALID, "too many rows for one launch");
    const bool vec = (F % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && (ldb % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                     (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    dim3 grid((unsigned)blocks), block(wpb * 32);
    if (!vec) {
        gather_rows_scalar_kernel<MODE><<<grid, block, 0, ctx->stream>>>(n_rows, F, rowptr, col, x, ldx, out, ldo, dinv, bias, relu, w_rows, ldb);
        GIGL_LAUNCHED(ctx);
        return GIGL_OK;
    }
    // heavy-row work list: a row needs > kRowSplit edges to get on it, so n_rows entries always suffice; capped at 1M
    const int32_t heavy_cap = (int32_t)(n_rows < (1 << 20) ? n_rows : (1 << 20));
    void* hl = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_WORK, sizeof(int32_t) * ((size_t)heavy_cap + 64), &hl);
    if (rc != GIGL_OK) return rc;
    GatherArgs a{n_rows, F, rowptr, col, x, ldx, out, ldo, dinv, bias, relu, w_rows, ldb, (int32_t*)hl, (int32_t*)hl + 64, heavy_cap};
    GIGL_CUDA(ctx, cudaMemsetAsync(a.heavy_ctr, 0, sizeof(int32_t), ctx->stream));
    if (F <= 16) {
        gather_rows_kernel<4, MODE><<<grid, block, 0, ctx->stream>>>(a);
    } else if (F <= 32) {
        gather_rows_kernel<8, MODE><<<grid, block, 0, ctx->stream>>>(a);
    } else if (F <= 64) {
        gather_rows_kernel<16, MODE><<<grid, block, 0, ctx->stream>>>(a);
    } else {
        gather_rows_kernel<32, MODE><<<grid, block, 0, ctx->stream>>>(a);
    }
    GIGL_LAUNCHED(ctx);
    gather_heavy_kernel<MODE><<<ctx->sm_count * 2, kHeavyWarpsAgg * 32, 0, ctx->stream>>>(a);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}

// carve a scratch slot into 256-byte aligned float arrays
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* b) : base((char*)b) {}
    float* take(size_t elems) {
        float* p = (float*)(base + off);
        off += (elems * sizeof(float) + 255) & ~(size_t)255;
        return p;
    }
    static size_t need(std::initializer_list<size_t> elems) {
        size_t t = 0;
        for (size_t e : elems) t += (e * sizeof(float) + 255) & ~(size_t)255;
        return t;
    }
};

}  // namespace gigl

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
    // inference: same kernels as the training forward ([mean | self] staged in library scratch, projection on tcgen05)
    void* a = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_SAVE, sizeof(float) * (size_t)n_rows_out * 2 * (size_t)((F + 3) & ~3), &a);
    if (rc != GIGL_OK) return rc;
    return sage_conv_train_fwd_launch(ctx, n, n_rows_out, F, O, rowptr, col, x, Wl, bl, Wr, out, (float*)a, relu);
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
    {   // x' = x W^T on the tensor cores (3xTF32)
        gigl_timed t(ctx, GIGL_T_GEMM_FULL);
        const int Fp = (F + 3) & ~3;
        const size_t x_el = (size_t)n * Fp, w_el = (size_t)O * Fp;
        void* hb = nullptr;
        if ((rc = gigl_scratch(ctx, GIGL_SLOT_SAVE, gigl::Carver::need({x_el, x_el, w_el, w_el}), &hb)) != GIGL_OK) return rc;
        gigl::Carver cv(hb);
        float *x_hi = cv.take(x_el), *x_lo = cv.take(x_el), *w_hi = cv.take(w_el), *w_lo = cv.take(w_el);
        if (Fp != F) GIGL_CUDA(ctx, cudaMemsetAsync(hb, 0, cv.off, ctx->stream));
        if ((rc = split_tf32_launch(ctx, n, F, x, F, x_hi, x_lo, Fp)) != GIGL_OK) return rc;
        if ((rc = split_tf32_launch(ctx, O, F, W, F, w_hi, w_lo, Fp)) != GIGL_OK) return rc;
        if ((rc = linear_tc_launch(ctx, n, O, Fp, x_hi, x_lo, Fp, w_hi, w_lo, Fp, nullptr, xp, O, 0)) != GIGL_OK) return rc;
    }
    const int wpb = 8;
    gigl::gcn_dinv_kernel<<<(unsigned)ceil_div64(n, wpb), wpb * 32, 0, ctx->stream>>>(n, rowptr, col, dinv);
    GIGL_LAUNCHED(ctx);
    return gigl::launch_gather<1>(ctx, n, O, rowptr, col, xp, out, dinv, b, relu);
}

// =============================================================================================
// Training forms: forward that keeps [mean | self] for the backward pass, and the backward pass.
// Reference call sites: out = model(x, edge_index); loss.backward()
//   python/gigl/src/common/modeling_task_specs/node_classification_modeling_task_spec.py:134-173
//   python/gigl/src/common/modeling_task_specs/graphsage_template_modeling_spec.py:299-367
// Projections run on tcgen05 (3xTF32): forward and grad_input through linear_tc_launch (K-major),
// weight gradients through linear_tn_tc_launch (reduction over the rows, MN-major operands).
// =============================================================================================
namespace gigl {

static inline int round4(int v) { return (v + 3) & ~3; }

// G = grad_out (masked by out > 0 when the layer fused a ReLU), pitch ldg, plus its TF32 hi / lo halves
__global__ void mask_split_kernel(int64_t rows, int cols, const float* __restrict__ grad, const float* __restrict__ out, int relu,
                                  float* __restrict__ G, float* __restrict__ hi, float* __restrict__ lo, int64_t ldg) {
    const int64_t total = rows * ldg;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ldg;
        const int c = (int)(i - r * ldg);
        float v = 0.f;
        if (c < cols) {
            v = grad[r * cols + c];
            if (relu && !(out[r * cols + c] > 0.f)) v = 0.f;
        }
        G[i] = v;
        gigl_split_tf32(v, hi[i], lo[i]);
    }
}

// [Wl | Wr] as one [O, 2 Fp] operand (zero padded), TF32 hi / lo
__global__ void wcat_split_kernel(int O, int F, int Fp, const float* __restrict__ Wl, const float* __restrict__ Wr,
                                  float* __restrict__ hi, float* __restrict__ lo) {
    const int total = O * 2 * Fp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int o = i / (2 * Fp), k = i - o * 2 * Fp;
        float v = 0.f;
        if (k < Fp) {
            if (k < F) v = Wl[(int64_t)o * F + k];
        } else if (Wr != nullptr && k - Fp < F) {
            v = Wr[(int64_t)o * F + (k - Fp)];
        }
        gigl_split_tf32(v, hi[i], lo[i]);
    }
}

// the transposed operand [2 Fp, Op]: row k = column k of [Wl | Wr]  (grad_input = G @ [Wl | Wr])
__global__ void wcat_t_split_kernel(int O, int Op, int F, int Fp, const float* __restrict__ Wl, const float* __restrict__ Wr,
                                    float* __restrict__ hi, float* __restrict__ lo) {
    const int rows = Wr != nullptr ? 2 * Fp : Fp;
    const int total = rows * Op;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / Op, o = i - k * Op;
        float v = 0.f;
        if (o < O) {
            if (k < Fp) {
                if (k < F) v = Wl[(int64_t)o * F + k];
            } else if (k - Fp < F) {
                v = Wr[(int64_t)o * F + (k - Fp)];
            }
        }
        gigl_split_tf32(v, hi[i], lo[i]);
    }
}

__global__ void inv_deg_kernel(int64_t m, const int64_t* __restrict__ rowptr, float* __restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t d = rowptr[i + 1] - rowptr[i];
    w[i] = 1.0f / (float)(d > 1 ? d : 1);
}

static inline unsigned grid_for(gigl_ctx* ctx, int64_t work, int per_block = 256) {
    int64_t g = ceil_div64(work > 0 ? work : 1, per_block);
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    return (unsigned)(g < cap ? g : cap);
}

}  // namespace gigl

int sage_conv_train_fwd_launch(gigl_ctx* ctx, int64_t n, int64_t m, int32_t F, int32_t O, const int64_t* rowptr, const int32_t* col,
                               const float* x, const float* Wl, const float* bl, const float* Wr, float* out, float* A_save,
                               int32_t relu) {
    using namespace gigl;
    GIGL_CHECK(ctx, n >= 0 && m >= 0 && m <= n && F >= 1 && O >= 1, "bad sizes");
    if (m == 0) return GIGL_OK;
    const int Fp = round4(F);
    const int64_t ldA = 2 * (int64_t)Fp;
    int rc;
    if (Fp != F) GIGL_CUDA(ctx, cudaMemsetAsync(A_save, 0, sizeof(float) * (size_t)m * ldA, ctx->stream));
    {
        gigl_timed t(ctx, GIGL_T_GATHER_FULL);
        if ((rc = launch_gather<0>(ctx, m, F, rowptr, col, x, A_save, nullptr, nullptr, 0, F, ldA)) != GIGL_OK) return rc;
        GIGL_CUDA(ctx, cudaMemcpy2DAsync(A_save + Fp, sizeof(float) * ldA, x, sizeof(float) * F, sizeof(float) * F, (size_t)m,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
    }
    gigl_timed t(ctx, GIGL_T_GEMM_FULL);
    const size_t a_el = (size_t)m * ldA, w_el = (size_t)O * ldA;
    void* buf = nullptr;
    if ((rc = gigl_scratch(ctx, GIGL_SLOT_AGG, Carver::need({a_el, a_el, w_el, w_el}), &buf)) != GIGL_OK) return rc;
    Carver cv(buf);
    float *a_hi = cv.take(a_el), *a_lo = cv.take(a_el), *w_hi = cv.take(w_el), *w_lo = cv.take(w_el);
    if ((rc = split_tf32_launch(ctx, m, (int)ldA, A_save, ldA, a_hi, a_lo, ldA)) != GIGL_OK) return rc;
    wcat_split_kernel<<<grid_for(ctx, (int64_t)w_el), 256, 0, ctx->stream>>>(O, F, Fp, Wl, Wr, w_hi, w_lo);
    GIGL_LAUNCHED(ctx);
    return linear_tc_launch(ctx, m, O, (int)ldA, a_hi, a_lo, ldA, w_hi, w_lo, ldA, bl, out, O, relu);
}

int sage_conv_bwd_launch(gigl_ctx* ctx, int64_t n, int64_t m, int32_t F, int32_t O, const int64_t* rowptr, const int64_t* t_rowptr,
                         const int32_t* t_col, const float* A_save, const float* Wl, const float* Wr, const float* out,
                         const float* grad_out, float* grad_x, float* grad_Wl, float* grad_bl, float* grad_Wr, int32_t relu) {
    using namespace gigl;
    GIGL_CHECK(ctx, n >= 0 && m >= 0 && m <= n && F >= 1 && O >= 1, "bad sizes");
    GIGL_CHECK(ctx, !relu || out != nullptr, "the fused ReLU needs the forward output for its mask");
    GIGL_CHECK(ctx, grad_x == nullptr || (t_rowptr != nullptr && t_col != nullptr), "grad_x needs the CSR by source");
    const int Fp = round4(F), Op = round4(O);
    const int64_t ldA = 2 * (int64_t)Fp;
    const size_t g_el = (size_t)m * Op, a_el = (size_t)m * ldA, wt_el = (size_t)ldA * Op;
    void* buf = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_AGG, Carver::need({g_el, g_el, g_el, a_el, a_el, wt_el, wt_el, (size_t)m}), &buf);
    if (rc != GIGL_OK) return rc;
    Carver cv(buf);
    float *G = cv.take(g_el), *g_hi = cv.take(g_el), *g_lo = cv.take(g_el);
    float *a_hi = cv.take(a_el), *a_lo = cv.take(a_el);  // [mean | self] halves; re-used for dA once the weight GEMM is done
    float *wt_hi = cv.take(wt_el), *wt_lo = cv.take(wt_el), *invdeg = cv.take((size_t)m);
    if (m > 0) {
        mask_split_kernel<<<grid_for(ctx, (int64_t)g_el), 256, 0, ctx->stream>>>(m, O, grad_out, out, relu, G, g_hi, g_lo, Op);
        GIGL_LAUNCHED(ctx);
    }
    if (grad_bl && (rc = colsum_launch(ctx, m, O, G, Op, grad_bl, 0)) != GIGL_OK) return rc;
    if (grad_Wl || grad_Wr) {
        GIGL_CHECK(ctx, grad_Wl && grad_Wr, "both weight gradients are produced together");
        gigl_timed t(ctx, GIGL_T_GEMM_FULL);
        if ((rc = split_tf32_launch(ctx, m, (int)ldA, A_save, ldA, a_hi, a_lo, ldA)) != GIGL_OK) return rc;
        if ((rc = linear_tn_tc_launch(ctx, m, O, (int)ldA, g_hi, g_lo, Op, a_hi, a_lo, ldA, grad_Wl, F, F, Fp, grad_Wr, F, F, 0)) != GIGL_OK)
            return rc;
    }
    if (grad_x) {
        float* dA = a_hi;  // [m, 2 Fp] = G @ [Wl | Wr]
        {
            gigl_timed t(ctx, GIGL_T_GEMM_FULL);
            wcat_t_split_kernel<<<grid_for(ctx, (int64_t)wt_el), 256, 0, ctx->stream>>>(O, Op, F, Fp, Wl, Wr, wt_hi, wt_lo);
            GIGL_LAUNCHED(ctx);
            if ((rc = linear_tc_launch(ctx, m, (int)ldA, Op, g_hi, g_lo, Op, wt_hi, wt_lo, Op, nullptr, dA, ldA, 0)) != GIGL_OK) return rc;
        }
        if (m > 0) {
            inv_deg_kernel<<<(unsigned)ceil_div64(m, 256), 256, 0, ctx->stream>>>(m, rowptr, invdeg);
            GIGL_LAUNCHED(ctx);
        }
        gigl_timed t(ctx, GIGL_T_GATHER_FULL);
        // grad_x[j] = sum_{i in OUT(j), i < m} dMean_i / deg_i + (j < m ? dSelf_j : 0)
        if ((rc = launch_gather<2>(ctx, n, F, t_rowptr, t_col, dA, grad_x, invdeg, dA + Fp, 0, ldA, F, m, ldA)) != GIGL_OK) return rc;
    }
    return GIGL_OK;
}

// C[M, N] = G[R, M]^T A[R, N]  (fp32 operands; split into TF32 halves here)
int linear_tn_launch(gigl_ctx* ctx, int64_t R, int M, int N, const float* G, int64_t ldg, const float* A, int64_t lda, float* C,
                     int64_t ldc, int accumulate) {
    using namespace gigl;
    if (R == 0) {  // empty reduction: C = 0 (or unchanged when accumulating)
        if (!accumulate) GIGL_CUDA(ctx, cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, (size_t)M, ctx->stream));
        return GIGL_OK;
    }
    const int Mp = round4(M), Np = round4(N);
    const size_t g_el = (size_t)R * Mp, a_el = (size_t)R * Np;
    void* buf = nullptr;
    int rc = gigl_scratch(ctx, GIGL_SLOT_AGG, Carver::need({g_el, g_el, a_el, a_el}), &buf);
    if (rc != GIGL_OK) return rc;
    Carver cv(buf);
    float *g_hi = cv.take(g_el), *g_lo = cv.take(g_el), *a_hi = cv.take(a_el), *a_lo = cv.take(a_el);
    if (Mp != M) {
        GIGL_CUDA(ctx, cudaMemsetAsync(g_hi, 0, sizeof(float) * g_el, ctx->stream));
        GIGL_CUDA(ctx, cudaMemsetAsync(g_lo, 0, sizeof(float) * g_el, ctx->stream));
    }
    if (Np != N) {
        GIGL_CUDA(ctx, cudaMemsetAsync(a_hi, 0, sizeof(float) * a_el, ctx->stream));
        GIGL_CUDA(ctx, cudaMemsetAsync(a_lo, 0, sizeof(float) * a_el, ctx->stream));
    }
    if ((rc = split_tf32_launch(ctx, R, M, G, ldg, g_hi, g_lo, Mp)) != GIGL_OK) return rc;
    if ((rc = split_tf32_launch(ctx, R, N, A, lda, a_hi, a_lo, Np)) != GIGL_OK) return rc;
    return linear_tn_tc_launch(ctx, R, M, N, g_hi, g_lo, Mp, a_hi, a_lo, Np, C, ldc, N, N, nullptr, 0, 0, accumulate);
}

// GCNConv backward.  out = Ahat (x W^T) + b with Ahat_ij = dinv_i dinv_j over the non-loop edges j -> i plus dinv_i^2 on the
// diagonal (dinv from the forward in-degrees), so d(xW^T) = Ahat^T g = the same weighted gather over the CSR by source.
int gcn_conv_bwd_launch(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr, const int32_t* col, const int64_t* t_rowptr,
                        const int32_t* t_col, const float* x, const float* W, const float* out, const float* grad_out, float* grad_x,
                        float* grad_W, float* grad_b, int32_t relu) {
    using namespace gigl;
    GIGL_CHECK(ctx, n >= 0 && F >= 1 && O >= 1 && t_rowptr && t_col, "bad arguments");
    GIGL_CHECK(ctx, !relu || out != nullptr, "the fused ReLU needs the forward output for its mask");
    if (n == 0) return GIGL_OK;
    const int Op = round4(O), Fp = round4(F);
    const size_t g_el = (size_t)n * Op;
    void* buf = nullptr;
    // SLOT_IO0: buffers that must survive the nested launches (which use SLOT_AGG / WORK / SORT)
    int rc = gigl_scratch(ctx, GIGL_SLOT_IO0, Carver::need({g_el, g_el, g_el, g_el, (size_t)n, (size_t)Fp * Op, (size_t)Fp * Op}), &buf);
    if (rc != GIGL_OK) return rc;
    Carver cv(buf);
    float *G = cv.take(g_el), *g_hi = cv.take(g_el), *g_lo = cv.take(g_el), *dxp = cv.take(g_el), *dinv = cv.take((size_t)n);
    float *wt_hi = cv.take((size_t)Fp * Op), *wt_lo = cv.take((size_t)Fp * Op);
    mask_split_kernel<<<grid_for(ctx, (int64_t)g_el), 256, 0, ctx->stream>>>(n, O, grad_out, out, relu, G, g_hi, g_lo, Op);
    GIGL_LAUNCHED(ctx);
    if (grad_b && (rc = colsum_launch(ctx, n, O, G, Op, grad_b, 0)) != GIGL_OK) return rc;
    const int wpb = 8;
    gcn_dinv_kernel<<<(unsigned)ceil_div64(n, wpb), wpb * 32, 0, ctx->stream>>>(n, rowptr, col, dinv);
    GIGL_LAUNCHED(ctx);
    // d(x W^T) [n, O] (pitch Op): MODE 1 over the transposed rows, no bias
    if (Op != O) GIGL_CUDA(ctx, cudaMemsetAsync(dxp, 0, sizeof(float) * g_el, ctx->stream));
    if ((rc = launch_gather<1>(ctx, n, O, t_rowptr, t_col, G, dxp, dinv, nullptr, 0, Op, Op)) != GIGL_OK) return rc;
    if (grad_W && (rc = linear_tn_launch(ctx, n, O, F, dxp, Op, x, F, grad_W, F, 0)) != GIGL_OK) return rc;
    if (grad_x) {
        // grad_x = d(xW^T) @ W: C[n, F] = dxp[n, O] @ (W^T)[F, O]^T
        wcat_t_split_kernel<<<grid_for(ctx, (int64_t)Fp * Op), 256, 0, ctx->stream>>>(O, Op, F, Fp, W, nullptr, wt_hi, wt_lo);
        GIGL_LAUNCHED(ctx);
        if ((rc = split_tf32_launch(ctx, n, Op, dxp, Op, g_hi, g_lo, Op)) != GIGL_OK) return rc;
        if ((rc = linear_tc_launch(ctx, n, F, Op, g_hi, g_lo, Op, wt_hi, wt_lo, Op, nullptr, grad_x, F, 0)) != GIGL_OK) return rc;
    }
    return GIGL_OK;
}
