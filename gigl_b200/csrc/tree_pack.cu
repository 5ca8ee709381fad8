// tree_pack.cu - the sampler's padded tree in its PACKED form for the trip to the host.
//
// gigl_sample_khop_* writes level h as n_roots * prod(fanouts[0..h]) slots with -1 padding (include/gigl_b200.h); on
// the products-like graph 61 % of the [15, 10] slots are filled.  The device-to-host copy of the index sets is the
// largest item of the host entry point (47.6 of 59.8 MB per 65 536 roots) and on an 8-GPU box the host side of PCIe,
// not the GPUs, bounds the end-to-end rate (measured: 8 ranks x 59.8 MB per 5.5 ms = 87 GB/s for the whole box).
// The packed form carries the same information in ~2/3 of the bytes: per parent slot its child count as ONE byte
// (fanout <= 128) and only the filled slots, hop after hop, parent slots in order, each parent's children in the
// order of the padded layout (a parent's filled slots are its first cnt slots).  This is the `Seq[Edge]` a
// KHopSamplerService hands back (scala_spark35/common/src/main/scala/graphdb/KHopSamplerService.scala:17-20): no padding.
#include <cuda_runtime.h>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace gigl {

__global__ void __launch_bounds__(256) pack_tree_kernel(int64_t n_slots, int32_t f, const int32_t* __restrict__ cnt,
                                                        const int32_t* __restrict__ goff, const int32_t* __restrict__ nbr,
                                                        int32_t* __restrict__ packed) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += stride) {
        const int64_t p = s / f;
        const int j = (int)(s - p * f);
        if (j < __ldg(cnt + p)) packed[__ldg(goff + p) + j] = __ldg(nbr + s);
    }
}

__global__ void __launch_bounds__(256) pack_counts_kernel(int64_t n_parents, const int32_t* __restrict__ cnt, uint8_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_parents; p += stride) out[p] = (uint8_t)__ldg(cnt + p);
}

// The filled slots again, `bits` bits each (ids < 2^bits), as one little-endian bit stream in 32-bit words: entry e lives in
// bits [e * bits, (e + 1) * bits).  One thread per OUTPUT word (it reads the two or three entries that overlap it): no atomics.
__global__ void __launch_bounds__(256) bitpack_kernel(const int32_t* __restrict__ n_dev, int bits, const int32_t* __restrict__ packed,
                                                      uint32_t* __restrict__ words) {
    const int64_t n = *n_dev;
    const int64_t n_words = (n * bits + 31) >> 5;
    const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_words; j += stride) {
        const int64_t bit0 = j << 5;
        int64_t e = bit0 / bits;
        int skip = (int)(bit0 - e * bits);  // low bits of entry e that went into the previous word
        uint32_t w = 0;
        for (int pos = 0; pos < 32 && e < n; ++e) {
            w |= (((uint32_t)__ldg(packed + e) & mask) >> skip) << pos;
            pos += bits - skip;
            skip = 0;
        }
        words[j] = w;
    }
}

}  // namespace gigl

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

// cnt_all_dev: the per-hop count arrays back to back (n_parents entries, one more readable behind them).
// Workspace: goff int32[n_parents + 1] | cnt_u8[n_parents] | packed int32[n_slots] | bit-packed words (id_bits > 0) | scan temp.
// goff[n_parents] = filled slots.
int tree_pack_launch(gigl_ctx* ctx, cudaStream_t st, int64_t n_roots, const int32_t* fanouts, int32_t n_hops,
                     const int32_t* const* nbr_dev, const int32_t* cnt_all_dev, int slot, int32_t** goff_dev, uint8_t** cnt_u8_dev,
                     int32_t** packed_dev, int id_bits, uint32_t** words_dev) {
    using namespace gigl;
    int64_t n_parents = 0, n_slots = 0, w = n_roots;
    for (int h = 0; h < n_hops; ++h) {
        n_parents += w;
        w *= fanouts[h];
        n_slots += w;
    }
    GIGL_CHECK(ctx, n_slots <= 0x7fffffffLL, "tree exceeds 2^31-1 slots; split the roots");
    size_t scan_bytes = 0;
    GIGL_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n_parents + 1), st));
    const size_t o_u8 = up256(sizeof(int32_t) * (size_t)(n_parents + 1));
    const size_t o_packed = o_u8 + up256((size_t)n_parents);
    const size_t o_words = o_packed + up256(sizeof(int32_t) * (size_t)(n_slots > 0 ? n_slots : 1));
    const size_t o_temp = o_words + (id_bits > 0 ? up256(sizeof(uint32_t) * (((size_t)n_slots * (size_t)id_bits + 31) / 32 + 1)) : 0);
    void* ws = nullptr;
    int rc = gigl_scratch(ctx, slot, o_temp + scan_bytes + 256, &ws);
    if (rc != GIGL_OK) return rc;
    int32_t* goff = (int32_t*)ws;
    uint8_t* u8 = (uint8_t*)ws + o_u8;
    int32_t* packed = (int32_t*)((char*)ws + o_packed);
    GIGL_CUDA(ctx, cub::DeviceScan::ExclusiveSum((char*)ws + o_temp, scan_bytes, cnt_all_dev, goff, (int)(n_parents + 1), st));
    ctx->launches++;
    const unsigned cap = (unsigned)ctx->sm_count * 16;
    int64_t p0 = 0;
    w = n_roots;
    for (int h = 0; h < n_hops; ++h) {
        const int64_t parents = w;
        w *= fanouts[h];
        if (w > 0) {
            const int64_t g = ceil_div64(w, 256);
            pack_tree_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(w, fanouts[h], cnt_all_dev + p0, goff + p0, nbr_dev[h], packed);
            GIGL_LAUNCHED(ctx);
        }
        p0 += parents;
    }
    if (n_parents > 0) {
        const int64_t g = ceil_div64(n_parents, 256);
        pack_counts_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(n_parents, cnt_all_dev, u8);
        GIGL_LAUNCHED(ctx);
    }
    if (id_bits > 0) {
        uint32_t* words = (uint32_t*)((char*)ws + o_words);
        const int64_t g = ceil_div64(ceil_div64(n_slots * id_bits, 32) + 1, 256);
        bitpack_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(goff + n_parents, id_bits, packed, words);
        GIGL_LAUNCHED(ctx);
        *words_dev = words;
    }
    *goff_dev = goff;
    *cnt_u8_dev = u8;
    *packed_dev = packed;
    return GIGL_OK;
}
