// frontier.cu - distinct frontier of a SamplingOp (typed / DAG sampling).
//
// The reference's graph-DB sampler expands, for every op, the SET of result nodes of the op's parents, once per node
// (GraphDBSampler.getKHopSubgraphForRootNode, scala_spark35/subgraph_sampler/src/main/scala/libs/sampler/GraphDBSampler.scala:66-82:
// the parents' results are collected into a HashSet[Node] before the op's query runs).  In the padded-tree layout a node
// reached along k paths occupies k slots of the parent level (possible from depth 2 on, and across the several inputs
// of one op), and expanding every slot would give that node up to k * fanout in-edges for one op where the reference
// bounds it by fanout.  This kernel writes a copy of the parent level in which, per root, only the FIRST slot of every
// distinct node survives (the others become -1 = "no group"), also against the slots of the op's earlier input
// instances; the sampler then expands that copy.
#include <cuda_runtime.h>

#include "common.cuh"

namespace gigl {

constexpr int kMaxPrevLists = 8;
struct PrevLists {
    const int32_t* ptr[kMaxPrevLists];
    int32_t slots[kMaxPrevLists];
    int32_t n;
};

// one warp per root: lane l owns slots l, l + 32, ...; a slot survives iff its node occurs in no earlier list and at no
// lower slot of its own row
__global__ void __launch_bounds__(256) frontier_distinct_kernel(int64_t n_roots, int32_t slots, const int32_t* __restrict__ cur,
                                                                const PrevLists prev, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_roots; r += warps) {
        const int32_t* row = cur + r * slots;
        for (int s = lane; s < slots; s += 32) {
            int32_t v = row[s];
            if (v >= 0) {
                bool seen = false;
                for (int t = 0; t < s && !seen; ++t) seen = row[t] == v;
                for (int p = 0; p < prev.n && !seen; ++p) {
                    const int32_t* pr = prev.ptr[p] + r * prev.slots[p];
                    for (int t = 0; t < prev.slots[p] && !seen; ++t) seen = pr[t] == v;
                }
                if (seen) v = -1;
            }
            out[r * slots + s] = v;
        }
    }
}

}  // namespace gigl

int gigl_frontier_distinct_dev(gigl_ctx* ctx, int64_t n_roots, int32_t n_prev, const int32_t* const* prev_dev,
                               const int32_t* prev_slots, const int32_t* cur_dev, int32_t cur_slots, int32_t* out_dev) {
    using namespace gigl;
    if (!ctx) return gigl_fail(nullptr, GIGL_E_INVALID, "null ctx");
    GIGL_CHECK(ctx, n_roots >= 0 && cur_slots >= 1 && n_prev >= 0 && n_prev <= kMaxPrevLists, "bad frontier arguments (at most 8 earlier lists)");
    GIGL_CHECK(ctx, (cur_dev && out_dev) || n_roots == 0, "null frontier");
    GIGL_CHECK(ctx, n_prev == 0 || (prev_dev && prev_slots), "null earlier lists");
    if (n_roots == 0) return GIGL_OK;
    GIGL_CUDA(ctx, cudaSetDevice(ctx->device));
    PrevLists pl{};
    pl.n = n_prev;
    for (int p = 0; p < n_prev; ++p) {
        GIGL_CHECK(ctx, prev_dev[p] && prev_slots[p] >= 1, "bad earlier list");
        pl.ptr[p] = prev_dev[p];
        pl.slots[p] = prev_slots[p];
    }
    const int64_t blocks = ceil_div64(n_roots, 8), cap = (int64_t)ctx->sm_count * 16;
    frontier_distinct_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, ctx->stream>>>(n_roots, cur_slots, cur_dev, pl, out_dev);
    GIGL_LAUNCHED(ctx);
    return GIGL_OK;
}
