/*
 * gigl_oracle.c - CPU restatement of the GiGL hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under gigl_b200/ may link, import or execute this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and there only as the checker / the CPU baseline - never as the product path.
 *
 * What it restates (paths relative to /root/reference):
 *   sampler  : scala/subgraph_sampler/src/main/scala/libs/task/SamplingStrategy.scala:16-82
 *              (hashBasedUniformPermutation), .../pureSpark/SGSPureSparkV1Task.scala:313-388
 *              (sampleOnehopSrcNodesUniformly), :390-494 (sampleTwohopSrcNodesUniformly),
 *              NodeAnchorBasedLinkPredictionBaseTask.scala:19-104 (sampleDstNodesUniformly).
 *   aggregate: python/gigl/src/common/models/pyg/homogeneous.py:107-202,488-546 and
 *              python/gigl/src/common/modeling_task_specs/graphsage_template_modeling_spec.py:143-148,
 *              i.e. torch_geometric 2.5.3 SAGEConv(mean)/GCNConv/GraphSAGE(BasicGNN) semantics
 *              (third-party, pinned python/pyproject.toml:53; source not in the tree, restated from
 *              its published definition - SURVEY.md Appendix B).
 *   third-party arithmetic: org.apache.spark:spark-sql_2.12:3.1.3 (scala/build.sbt:37-38)
 *              XXH64.hashInt(int, seed=42L), array_sort on struct<_hash:bigint,_indices:int>
 *              (ascending, signed, lexicographic), non-ANSI IntegerType '+' (32-bit wrap).
 *
 * Pinning status: the reference's own tests never run the deterministic permutation
 * (every call passes permutationStrategy="non-deterministic",
 * scala/subgraph_sampler/src/test/scala/SGSPureSparkV1TaskTest.scala:200,225,256,288), so
 * there is no reference golden for it: **parity unpinned** at that level.  What IS pinned
 * (tests/test_oracle_*.py): the hash against Spark's documented xxhash64 KAT and an
 * independent XXH64 implementation (tests/golden/xxh64_kat.json); the structural rules
 * against the reference sampler's real outputs (tests/golden/ *_sgs_output.json), exact for
 * every root whose frontier degrees are all <= fanout.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <omp.h>

#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_3 0x165667B19E3779F9ULL
#define P64_5 0x27D4EB2F165667C5ULL

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

/* Spark XXH64.hashInt(int input, long seed) == XXH64 of the 4 little-endian bytes. */
int64_t oracle_xxh64_int(int32_t input, uint64_t seed) {
    uint64_t h = seed + P64_5 + 4ULL;
    h ^= (uint64_t)(uint32_t)input * P64_1;
    h = rotl64(h, 23) * P64_2 + P64_3;
    h ^= h >> 33;
    h *= P64_2;
    h ^= h >> 29;
    h *= P64_3;
    h ^= h >> 32;
    return (int64_t)h;
}

/* int32 add with two's-complement wrap (Spark IntegerType '+', ANSI off). */
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

typedef struct {
    int64_t key;
    int32_t idx; /* 1-based, as F.sequence(1, size) */
} kv_t;

static int kv_cmp(const void* a, const void* b) {
    const kv_t* x = (const kv_t*)a;
    const kv_t* y = (const kv_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
    return 0;
}

/*
 * Full permutation, literally as SamplingStrategy.scala:47-76: hash every index 1..size with
 * x + _internal_seed + currentSeed, array_sort the (hash, index) structs, emit indices (0-based here).
 */
int oracle_perm_full(int64_t size, int32_t internal_seed, int32_t current_seed, int64_t* out_idx0) {
    kv_t* kv = (kv_t*)malloc(sizeof(kv_t) * (size_t)(size > 0 ? size : 1));
    if (!kv) return -1;
    for (int64_t i = 1; i <= size; ++i) {
        int32_t x = wadd(wadd((int32_t)i, internal_seed), current_seed);
        kv[i - 1].key = oracle_xxh64_int(x, 42ULL);
        kv[i - 1].idx = (int32_t)i;
    }
    qsort(kv, (size_t)size, sizeof(kv_t), kv_cmp);
    for (int64_t i = 0; i < size; ++i) out_idx0[i] = kv[i].idx - 1;
    free(kv);
    return 0;
}

/*
 * First `f` entries of the same permutation (slice(_shuffled, 1, f)) without sorting everything:
 * bounded max-heap on (key, idx).  Returns the number written = min(f, size); out_idx0 is in
 * permutation order (ascending key).
 */
static inline int kv_less(kv_t a, kv_t b) { return a.key < b.key || (a.key == b.key && a.idx < b.idx); }

static int64_t perm_topk(int64_t size, int32_t internal_seed, int32_t current_seed, int32_t f, kv_t* heap,
                         int64_t* out_idx0) {
    int64_t n = 0;
    const int32_t base = wadd(internal_seed, current_seed);
    for (int64_t i = 1; i <= size; ++i) {
        kv_t c;
        c.key = oracle_xxh64_int(wadd((int32_t)i, base), 42ULL);
        c.idx = (int32_t)i;
        if (n < f) { /* sift up (max-heap) */
            int64_t p = n++;
            heap[p] = c;
            while (p > 0) {
                int64_t q = (p - 1) / 2;
                if (kv_less(heap[q], heap[p])) {
                    kv_t t = heap[q];
                    heap[q] = heap[p];
                    heap[p] = t;
                    p = q;
                } else
                    break;
            }
        } else if (kv_less(c, heap[0])) {
            int64_t p = 0;
            heap[0] = c;
            for (;;) {
                int64_t l = 2 * p + 1, r = l + 1, m = p;
                if (l < n && kv_less(heap[m], heap[l])) m = l;
                if (r < n && kv_less(heap[m], heap[r])) m = r;
                if (m == p) break;
                kv_t t = heap[m];
                heap[m] = heap[p];
                heap[p] = t;
                p = m;
            }
        }
    }
    qsort(heap, (size_t)n, sizeof(kv_t), kv_cmp);
    for (int64_t i = 0; i < n; ++i) out_idx0[i] = heap[i].idx - 1;
    return n;
}

int64_t oracle_perm_topk(int64_t size, int32_t internal_seed, int32_t current_seed, int32_t f, int64_t* out_idx0) {
    kv_t* heap = (kv_t*)malloc(sizeof(kv_t) * (size_t)(f > 0 ? f : 1));
    int64_t n = perm_topk(size, internal_seed, current_seed, f, heap, out_idx0);
    free(heap);
    return n;
}

/*
 * k-hop rooted-neighbourhood index sampling over a CSR whose row v lists the sorted
 * (ascending, duplicates kept) in-neighbours of v  [array_sort(collect_list(_src_node)),
 * SGSPureSparkV1Task.scala:334-342].
 *
 * Hop h (1-based) uses fanouts[h-1], currentSeed = base_seed * (first_call_no + h - 1)
 * [SamplingStrategy.scala:36,79] and _internal_seed = sum of the path ids root..hop h-1
 * [:39-45; `_dst_node` at hop 1, `_0_hop+_1_hop` at hop 2].  With n_hops = 2 and equal fanouts
 * this is exactly the reference; other fanout lists are the per-hop generalisation in
 * SURVEY.md section 8(a).
 *
 * GROUP BY (_0_hop,_1_hop) semantics [:438-448]: if the sampled list of one parent holds the
 * same vertex k m>1 times (only possible with duplicate directed edges), explode+join yields
 * ONE group whose array is IN(k) with every element repeated m times, i.e. size m*deg(k) and
 * sorted position j -> IN(k)[j / m].  The group is stored at the FIRST slot holding k; the
 * other duplicate slots get count 0.
 *
 * Layout (the "padded tree"): level h has n_roots * prod(fanouts[0..h-1]) int32 slots in
 * nbr[h-1], -1 = empty; slot s at level h is child (s % f_h) of slot (s / f_h) at level h-1
 * (level 0 = roots).  cnt[h-1][parent_slot] = number of children written.
 * Rows longer than 2^31-1 after multiplicity are rejected (-2): F.sequence is int-indexed.
 */
int oracle_sample_khop(int64_t n_nodes, const int64_t* rowptr, const int32_t* col, const int32_t* roots,
                       int64_t n_roots, const int32_t* fanouts, int32_t n_hops, int32_t base_seed,
                       int32_t first_call_no, int32_t** nbr /* [n_hops] */, int32_t** cnt /* [n_hops] */,
                       int32_t n_threads) {
    (void)n_threads;
    if (n_hops < 1 || n_hops > 8) return -1;
    int64_t width_prev = 1; /* slots per root at level h-1 */
    int rc = 0;
    for (int32_t h = 1; h <= n_hops; ++h) {
        const int32_t f = fanouts[h - 1];
        if (f < 1) return -1;
        const int32_t cur_seed = wmul(base_seed, wadd(first_call_no, h - 1));
        const int64_t width = width_prev * f;
        int32_t* out = nbr[h - 1];
        int32_t* oc = cnt[h - 1];
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
        {
            kv_t* heap = (kv_t*)malloc(sizeof(kv_t) * (size_t)f);
            int64_t* sel = (int64_t*)malloc(sizeof(int64_t) * (size_t)f);
#pragma omp for schedule(dynamic, 64)
            for (int64_t r = 0; r < n_roots; ++r) {
                for (int64_t ps = 0; ps < width_prev; ++ps) { /* parent slot within root */
                    const int64_t pslot = r * width_prev + ps;
                    int32_t* o = out + pslot * f;
                    for (int32_t j = 0; j < f; ++j) o[j] = -1;
                    oc[pslot] = 0;
                    /* parent vertex + path-id sum */
                    int32_t v, ssum;
                    if (h == 1) {
                        v = roots[r];
                        ssum = v;
                    } else {
                        v = nbr[h - 2][pslot];
                        if (v < 0) continue;
                        /* walk up the tree to sum path ids */
                        ssum = v;
                        int64_t s = pslot;
                        for (int32_t hh = h - 1; hh >= 1; --hh) {
                            s /= fanouts[hh - 1];
                            ssum = wadd(ssum, hh == 1 ? roots[s] : nbr[hh - 2][s]);
                        }
                        /* multiplicity among the siblings (same parent) */
                        const int64_t sib0 = (pslot / fanouts[h - 2]) * fanouts[h - 2];
                        int32_t m = 0;
                        int first = 1;
                        for (int32_t j = 0; j < fanouts[h - 2]; ++j) {
                            if (nbr[h - 2][sib0 + j] == v) {
                                if (sib0 + j < pslot) first = 0;
                                ++m;
                            }
                        }
                        if (!first) continue;
                        if (v >= n_nodes) {
#pragma omp atomic write
                            rc = -3;
                            continue;
                        }
                        const int64_t d = rowptr[v + 1] - rowptr[v];
                        const int64_t size = d * m;
                        if (size > 2147483647LL) {
#pragma omp atomic write
                            rc = -2;
                            continue;
                        }
                        int64_t n = perm_topk(size, ssum, cur_seed, f, heap, sel);
                        for (int64_t j = 0; j < n; ++j) o[j] = col[rowptr[v] + sel[j] / m];
                        oc[pslot] = (int32_t)n;
                        continue;
                    }
                    if (v < 0 || v >= n_nodes) {
#pragma omp atomic write
                        rc = -3;
                        continue;
                    }
                    const int64_t d = rowptr[v + 1] - rowptr[v];
                    if (d > 2147483647LL) {
#pragma omp atomic write
                        rc = -2;
                        continue;
                    }
                    int64_t n = perm_topk(d, ssum, cur_seed, f, heap, sel);
                    for (int64_t j = 0; j < n; ++j) o[j] = col[rowptr[v] + sel[j]];
                    oc[pslot] = (int32_t)n;
                }
            }
            free(heap);
            free(sel);
        }
        width_prev = width;
    }
    return rc;
}

/*
 * Positive-edge sampling (NodeAnchorBasedLinkPredictionBaseTask.scala:19-104): same permutation
 * over OUT(u) (CSR by src, sorted dst), internal seed = u, currentSeed = base_seed * call_no
 * (call_no = 3 in the NABLP task: hop1, hop2, positives).  It is oracle_sample_khop with
 * n_hops = 1 on the out-CSR and first_call_no = call_no; provided for readability.
 */
int oracle_sample_positives(int64_t n_nodes, const int64_t* out_rowptr, const int32_t* out_col,
                            const int32_t* srcs, int64_t n_srcs, int32_t num_pos, int32_t base_seed,
                            int32_t call_no, int32_t* pos, int32_t* pos_cnt, int32_t n_threads) {
    int32_t* nbr[1] = {pos};
    int32_t* cnt[1] = {pos_cnt};
    return oracle_sample_khop(n_nodes, out_rowptr, out_col, srcs, n_srcs, &num_pos, 1, base_seed, call_no, nbr,
                              cnt, n_threads);
}

/* ------------------------------------------------------------------------------------------
 * Aggregate half.  COO input exactly like PyG: edge_index[0] = src j, edge_index[1] = dst i.
 * ------------------------------------------------------------------------------------------ */

/*
 * SAGEConv(aggr=mean, root_weight=True, bias=True, normalize=False, project=False):
 *   m_i  = (1 / max(1, |{e: dst_e = i}|)) * sum_{e: dst_e = i} x[src_e]
 *   out_i = lin_l.weight @ m_i + lin_l.bias + lin_r.weight @ x_i        [SURVEY.md Appendix B]
 * fp32 throughout, edges accumulated in input order (index_add_ on CPU), F.linear as a plain
 * sequential dot product.  relu != 0 applies the inter-layer ReLU of BasicGNN.
 */
int oracle_sage_conv_f32(int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* src, const int64_t* dst,
                         const float* x, const float* Wl, const float* bl, const float* Wr, float* out,
                         int32_t relu, int32_t n_threads) {
    float* agg = (float*)calloc((size_t)(n * F > 0 ? n * F : 1), sizeof(float));
    int32_t* c = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    if (!agg || !c) return -1;
    for (int64_t k = 0; k < e; ++k) {
        if (src[k] < 0 || src[k] >= n || dst[k] < 0 || dst[k] >= n) {
            free(agg);
            free(c);
            return -3;
        }
        float* a = agg + dst[k] * F;
        const float* xs = x + src[k] * F;
        for (int32_t j = 0; j < F; ++j) a[j] += xs[j];
        c[dst[k]]++;
    }
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float inv = 1.0f / (float)(c[i] > 1 ? c[i] : 1);
        float* a = agg + i * F;
        for (int32_t j = 0; j < F; ++j) a[j] = a[j] * inv;
        const float* xi = x + i * F;
        for (int32_t o = 0; o < O; ++o) {
            float s = 0.f, t = 0.f;
            const float* wl = Wl + (int64_t)o * F;
            const float* wr = Wr + (int64_t)o * F;
            for (int32_t j = 0; j < F; ++j) s += wl[j] * a[j];
            for (int32_t j = 0; j < F; ++j) t += wr[j] * xi[j];
            float v = s + (bl ? bl[o] : 0.f) + t;
            out[i * O + o] = (relu && v < 0.f) ? 0.f : v;
        }
    }
    free(agg);
    free(c);
    return 0;
}

/* Same op in fp64 (inputs fp32, everything accumulated in double): the tolerance reference. */
int oracle_sage_conv_f64(int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* src, const int64_t* dst,
                         const float* x, const float* Wl, const float* bl, const float* Wr, double* out,
                         int32_t relu, int32_t n_threads) {
    double* agg = (double*)calloc((size_t)(n * F > 0 ? n * F : 1), sizeof(double));
    int32_t* c = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    if (!agg || !c) return -1;
    for (int64_t k = 0; k < e; ++k) {
        double* a = agg + dst[k] * F;
        const float* xs = x + src[k] * F;
        for (int32_t j = 0; j < F; ++j) a[j] += (double)xs[j];
        c[dst[k]]++;
    }
#pragma omp parallel for num_threads(n_threads > 0 ? n_threads : 1) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const double inv = 1.0 / (double)(c[i] > 1 ? c[i] : 1);
        double* a = agg + i * F;
        const float* xi = x + i * F;
        for (int32_t o = 0; o < O; ++o) {
            double s = 0.0;
            const float* wl = Wl + (int64_t)o * F;
            const float* wr = Wr + (int64_t)o * F;
            for (int32_t j = 0; j < F; ++j) s += (double)wl[j] * (a[j] * inv) + (double)wr[j] * (double)xi[j];
            s += bl ? (double)bl[o] : 0.0;
            out[i * O + o] = (relu && s < 0.0) ? 0.0 : s;
        }
    }
    free(agg);
    free(c);
    return 0;
}

/*
 * GCNConv(F, O) with add_self_loops=True, normalize=True, edge weights 1 (gcn_norm +
 * add_remaining_self_loops): x' = x @ lin.weight^T; every node WITHOUT a self loop gets one
 * (weight 1; existing self loops are kept as they are); deg_i = sum_{e: dst_e = i} w_e;
 * out_i = sum_{e: dst_e = i} deg_{src_e}^{-1/2} deg_i^{-1/2} x'_{src_e} + bias.
 * [SURVEY.md Appendix B; reference use: homogeneous.py:527-539].
 * add_remaining_self_loops keeps non-loop edges in order, then appends one loop per node (the
 * existing loop's weight if any - here always 1, but duplicates of a self loop collapse to ONE).
 */
#define GCN_BODY(T)                                                                                            \
    T* xp = (T*)calloc((size_t)(n * O > 0 ? n * O : 1), sizeof(T));                                            \
    T* deg = (T*)calloc((size_t)(n > 0 ? n : 1), sizeof(T));                                                   \
    if (!xp || !deg) return -1;                                                                                \
    for (int64_t i = 0; i < n; ++i)                                                                            \
        for (int32_t o = 0; o < O; ++o) {                                                                      \
            T s = 0;                                                                                           \
            for (int32_t j = 0; j < F; ++j) s += (T)W[(int64_t)o * F + j] * (T)x[i * F + j];                   \
            xp[i * O + o] = s;                                                                                 \
        }                                                                                                      \
    for (int64_t k = 0; k < e; ++k)                                                                            \
        if (src[k] != dst[k]) deg[dst[k]] += 1;                                                                \
    for (int64_t i = 0; i < n; ++i) deg[i] += 1; /* exactly one self loop per node */                          \
    for (int64_t i = 0; i < n * O; ++i) out[i] = 0;                                                            \
    for (int64_t k = 0; k < e; ++k) {                                                                          \
        if (src[k] == dst[k]) continue;                                                                        \
        T w = ((T)1 / (T)SQRT(deg[src[k]])) * ((T)1 / (T)SQRT(deg[dst[k]]));                                   \
        for (int32_t o = 0; o < O; ++o) out[dst[k] * O + o] += w * xp[src[k] * O + o];                         \
    }                                                                                                          \
    for (int64_t i = 0; i < n; ++i) {                                                                          \
        T w = ((T)1 / (T)SQRT(deg[i])) * ((T)1 / (T)SQRT(deg[i]));                                             \
        for (int32_t o = 0; o < O; ++o) {                                                                      \
            T v = out[i * O + o] + w * xp[i * O + o] + (b ? (T)b[o] : (T)0);                                   \
            out[i * O + o] = (relu && v < 0) ? (T)0 : v;                                                       \
        }                                                                                                      \
    }                                                                                                          \
    free(xp);                                                                                                  \
    free(deg);                                                                                                 \
    return 0;

int oracle_gcn_conv_f32(int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* src, const int64_t* dst,
                        const float* x, const float* W, const float* b, float* out, int32_t relu) {
#define SQRT sqrtf
    GCN_BODY(float)
#undef SQRT
}

int oracle_gcn_conv_f64(int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* src, const int64_t* dst,
                        const float* x, const float* W, const float* b, double* out, int32_t relu) {
#define SQRT sqrt
    GCN_BODY(double)
#undef SQRT
}

/* ---------------------------------------------------------------------------------------------
 * Batch collation for the TIMED CPU baseline (bench.py): the union of B sampled subgraphs with nodes de-duplicated by
 * id and edges by (src, dst), as the reference's GraphBuilder / collate fns build it
 * (python/gigl/src/common/graph_builder/abstract_graph_builder.py:49-197, pyg_graph_builder.py:20-69,
 * training/v1/lib/data_loaders/rooted_node_neighborhood_data_loader.py:78-) - the same result as
 * oracle.np_collate_fast (unique roots in first-occurrence order, then the other nodes ascending; edges ascending by
 * (dst, src) global id), with every core of the host: a parallel LSD radix sort of the edge keys and dense per-vertex
 * maps.  Returns 0, or -1 on a bad argument / allocation failure.
 *   nbr[h]: int32 [n_roots * prod fanouts[0..h]] padded tree (-1 = empty slot), level h + 1 children of level h
 *   node_ids (cap n_roots + slots), edge_src / edge_dst (cap slots, local ids), root_index [n_roots]            */
static void radix_pass_u64(const uint64_t* in, uint64_t* out, int64_t n, int shift, int n_threads, int64_t* hist /* [T][256] */) {
#pragma omp parallel num_threads(n_threads)
    {
        const int t = omp_get_thread_num(), nt = omp_get_num_threads();  /* the team may be smaller than asked for */
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        int64_t* h = hist + (int64_t)t * 256;
        for (int d = 0; d < 256; ++d) h[d] = 0;
        for (int64_t i = lo; i < hi; ++i) h[(in[i] >> shift) & 255]++;
#pragma omp barrier
#pragma omp single
        {
            int64_t run = 0;
            for (int d = 0; d < 256; ++d)
                for (int tt = 0; tt < nt; ++tt) {
                    const int64_t c = hist[(int64_t)tt * 256 + d];
                    hist[(int64_t)tt * 256 + d] = run;
                    run += c;
                }
        }
        for (int64_t i = lo; i < hi; ++i) out[h[(in[i] >> shift) & 255]++] = in[i];
    }
}

int oracle_collate(int64_t n_graph_nodes, const int32_t* roots, int64_t n_roots, const int32_t* fanouts, int32_t n_hops,
                   const int32_t* const* nbr, int64_t* node_ids, int64_t* n_nodes_out, int64_t* edge_src, int64_t* edge_dst,
                   int64_t* n_edges_out, int64_t* root_index, int32_t n_threads) {
    if (n_graph_nodes <= 0 || n_roots < 0 || n_hops < 1 || !fanouts || !nbr) return -1;
    if (n_threads <= 0) n_threads = omp_get_max_threads();
    int64_t slots = 0, width = n_roots;
    for (int h = 0; h < n_hops; ++h) {
        width *= fanouts[h];
        slots += width;
    }
    uint64_t* ka = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(slots > 0 ? slots : 1));
    uint64_t* kb = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(slots > 0 ? slots : 1));
    int64_t* hist = (int64_t*)malloc(sizeof(int64_t) * 256 * (size_t)n_threads);
    int32_t* lid = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_graph_nodes);
    if (!ka || !kb || !hist || !lid) {
        free(ka); free(kb); free(hist); free(lid);
        return -1;
    }
    /* 1. keys dst << 32 | src of the filled slots */
    int64_t n_keys = 0;
    width = n_roots;
    for (int h = 0; h < n_hops; ++h) {
        const int32_t* parents = h == 0 ? roots : nbr[h - 1];
        const int32_t f = fanouts[h];
        width *= f;
        for (int64_t s = 0; s < width; ++s) {
            const int32_t c = nbr[h][s];
            if (c >= 0) ka[n_keys++] = ((uint64_t)(uint32_t)parents[s / f] << 32) | (uint32_t)c;
        }
    }
    /* 2. sort (bytes that can be non-zero only), unique */
    int id_bits = 1;
    while (id_bits < 32 && ((int64_t)1 << id_bits) < n_graph_nodes) ++id_bits;
    uint64_t *in = ka, *out = kb;
    for (int part = 0; part < 2; ++part)
        for (int b = 0; b * 8 < id_bits; ++b) {
            radix_pass_u64(in, out, n_keys, part * 32 + b * 8, n_threads, hist);
            uint64_t* t = in; in = out; out = t;
        }
    int64_t n_edges = 0;
    for (int64_t i = 0; i < n_keys; ++i)
        if (i == 0 || in[i] != in[i - 1]) out[n_edges++] = in[i];
    const uint64_t* uk = out;
    /* 3. nodes: unique roots in first-occurrence order, then every other endpoint ascending (dense maps) */
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int64_t v = 0; v < n_graph_nodes; ++v) lid[v] = -1;
    int64_t n_nodes = 0;
    for (int64_t i = 0; i < n_roots; ++i) {
        const int32_t r = roots[i];
        if (r < 0 || r >= n_graph_nodes) { free(ka); free(kb); free(hist); free(lid); return -1; }
        if (lid[r] < 0) {
            lid[r] = (int32_t)n_nodes;
            node_ids[n_nodes++] = r;
        }
    }
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int64_t i = 0; i < n_edges; ++i) {  /* -2 = endpoint that is not a root (racing writers store the same value) */
        const int32_t s = (int32_t)(uk[i] & 0xffffffffu), d = (int32_t)(uk[i] >> 32);
        int32_t cur;
#pragma omp atomic read
        cur = lid[s];
        if (cur == -1) {
#pragma omp atomic write
            lid[s] = -2;
        }
#pragma omp atomic read
        cur = lid[d];
        if (cur == -1) {
#pragma omp atomic write
            lid[d] = -2;
        }
    }
    {
        int64_t* cnt = hist;  /* reuse: per-thread counts */
        for (int t = 0; t < n_threads; ++t) cnt[t] = 0;
#pragma omp parallel num_threads(n_threads)
        {
            const int t = omp_get_thread_num(), nt = omp_get_num_threads();
            const int64_t lo = n_graph_nodes * t / nt, hi = n_graph_nodes * (t + 1) / nt;
            int64_t c = 0;
            for (int64_t v = lo; v < hi; ++v) c += lid[v] == -2;
            cnt[t] = c;
#pragma omp barrier
            int64_t base = n_nodes;
            for (int tt = 0; tt < t; ++tt) base += cnt[tt];
            for (int64_t v = lo; v < hi; ++v)
                if (lid[v] == -2) {
                    lid[v] = (int32_t)base;
                    node_ids[base++] = v;
                }
        }
        for (int t = 0; t < n_threads; ++t) n_nodes += cnt[t];
    }
    /* 4. edge_index in local ids, root rows */
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int64_t i = 0; i < n_edges; ++i) {
        edge_src[i] = lid[(int32_t)(uk[i] & 0xffffffffu)];
        edge_dst[i] = lid[(int32_t)(uk[i] >> 32)];
    }
    for (int64_t i = 0; i < n_roots; ++i) root_index[i] = lid[roots[i]];
    *n_nodes_out = n_nodes;
    *n_edges_out = n_edges;
    free(ka); free(kb); free(hist); free(lid);
    return 0;
}
