/*
 * gigl_b200.h - C-ABI of libgigl_b200.so: the B200 (sm_100a) implementation of GiGL's two
 * data-parallel hot paths, k-hop rooted-neighbourhood sampling and the per-layer GNN
 * message-passing aggregate.  Plain C: pointers and sizes only, no C++/torch types, no
 * exceptions across the boundary.
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, < 0 = error (GIGL_E_*); the message is
 *     available from gigl_last_error(ctx) until the next call on that ctx;
 *   - one gigl_ctx per GPU (device + one CUDA stream + scratch); calls on one ctx are NOT
 *     thread-safe, distinct ctxs may be used from distinct threads (this matches the reference's
 *     per-partition setup()/teardown() of a sampler service,
 *     scala_spark35/common/src/main/scala/graphdb/KHopSamplerService.scala:10-33);
 *   - the caller owns every host buffer; the library owns device memory behind the opaque handles;
 *   - "*_host" entry points take host pointers and do the host<->device copies themselves
 *     (what a JNI / ctypes binding calls); "*_dev" entry points take device pointers, enqueue on
 *     the ctx stream and do not synchronise (what the PyTorch extension calls);
 *   - there is NO CPU fallback: without a usable CUDA device every entry point fails with
 *     GIGL_E_CUDA.
 *
 * Reference interfaces replaced (paths relative to the GiGL repository root):
 *   gigl_graph_*            <- loadEdgeDataframeIntoSparkSql / loadUnhydratedEdgeDataframeIntoSparkSql
 *                              scala/subgraph_sampler/src/main/scala/libs/task/pureSpark/SGSPureSparkV1Task.scala:120-311
 *   gigl_sample_khop_*      <- sampleOnehopSrcNodesUniformly :313-388 + sampleTwohopSrcNodesUniformly :390-494
 *                              with SamplingStrategy.hashBasedUniformPermutation
 *                              scala/subgraph_sampler/src/main/scala/libs/task/SamplingStrategy.scala:16-82;
 *                              plugin-side equivalent KHopSamplerService.getKHopSubgraphForRootNodes
 *                              scala_spark35/common/src/main/scala/graphdb/KHopSamplerService.scala:17-20
 *   gigl_sample_positives_* <- sampleDstNodesUniformly
 *                              scala/subgraph_sampler/src/main/scala/libs/task/pureSpark/NodeAnchorBasedLinkPredictionBaseTask.scala:19-104
 *   gigl_sage_conv_*        <- torch_geometric.nn.SAGEConv.forward as built by GraphSAGE.init_conv_layers
 *                              python/gigl/src/common/models/pyg/homogeneous.py:171-202 and called at
 *                              python/gigl/src/common/modeling_task_specs/graphsage_template_modeling_spec.py:305-311
 *   gigl_gcn_conv_*         <- torch_geometric.nn.GCNConv.forward used by TwoLayerGCN
 *                              python/gigl/src/common/models/pyg/homogeneous.py:527-542
 *   gigl_csr_from_coo_dev   <- the (src,dst) edge_index convention of PygGraphBuilder.build
 *                              python/gigl/src/common/graph_builder/pyg_graph_builder.py:20-69
 */
#ifndef GIGL_B200_H_
#define GIGL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIGL_OK 0
#define GIGL_E_INVALID -1   /* bad argument (null pointer, negative size, fanout out of range ...) */
#define GIGL_E_CUDA -2      /* CUDA runtime / driver error, or no usable device */
#define GIGL_E_RANGE -3     /* a vertex id outside [0, n_nodes) was met on the device */
#define GIGL_E_OVERFLOW -4  /* a row (times its multiplicity) exceeds 2^31-1 entries */
#define GIGL_E_NOMEM -5     /* device or host allocation failed */

#define GIGL_MAX_HOPS 8
#define GIGL_MAX_FANOUT 128

typedef struct gigl_ctx gigl_ctx;
typedef struct gigl_graph gigl_graph;

/* ---- library / context ------------------------------------------------------------------ */

/* "gigl_b200 <version> sm_100a" */
const char* gigl_version(void);

/* Creates a context on `device` with its own non-blocking stream. */
int gigl_ctx_create(int device, gigl_ctx** out);
/* Same, but enqueues on a stream owned by the caller (e.g. torch's current stream, passed as
 * the cudaStream_t value).  The stream must outlive the ctx. */
int gigl_ctx_create_on_stream(int device, void* cuda_stream, gigl_ctx** out);
void gigl_ctx_destroy(gigl_ctx* ctx);
/* Blocks until everything enqueued on the ctx stream is done; reports deferred device errors
 * (GIGL_E_RANGE / GIGL_E_OVERFLOW raised by the last *_dev sampling call). */
int gigl_ctx_sync(gigl_ctx* ctx);
const char* gigl_last_error(gigl_ctx* ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
int64_t gigl_ctx_launch_count(gigl_ctx* ctx);
/* The cudaStream_t the ctx enqueues on (for event timing on the launching stream). */
void* gigl_ctx_stream(gigl_ctx* ctx);
/*
 * Phase timing: when enabled, the library brackets its kernel groups with CUDA events on the ctx
 * stream and accumulates device milliseconds per tag ("sample", "collate_sort", "gather_l1",
 * "gemm_l1", ...; see gigl_timing_tag_name).  This is the hook the reference's coarse wall timers
 * (gigl/common/metrics/decorators.py:68-111) and TorchProfiler wrapper occupy; bench.py reads its
 * per-kernel roofline from it.  gigl_ctx_get_timing synchronises the stream.
 */
int gigl_ctx_set_timing(gigl_ctx* ctx, int32_t enabled);
int gigl_ctx_reset_timing(gigl_ctx* ctx);
int gigl_ctx_get_timing(gigl_ctx* ctx, int32_t tag, double* total_ms, int64_t* count);
int32_t gigl_timing_num_tags(void);
const char* gigl_timing_tag_name(int32_t tag);

/* ---- graph: CSR by destination, resident in HBM ---------------------------------------- */

/*
 * rowptr: int64[n_nodes + 1]; col: int32[n_edges], row v = the in-neighbours (sources) of v in
 * ASCENDING order, duplicates kept (array_sort(collect_list(_src_node)),
 * SGSPureSparkV1Task.scala:334-342).  Copies both arrays to the device.
 */
int gigl_graph_create_host(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int64_t* rowptr,
                           const int32_t* col, gigl_graph** out);
/* Wraps device arrays already laid out as above; nothing is copied, the caller keeps ownership. */
int gigl_graph_wrap_dev(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int64_t* rowptr_dev,
                        const int32_t* col_dev, gigl_graph** out);
/*
 * Builds the sorted in-CSR on the device from a host edge list, applying the reference's load
 * rules: ids are int32; if !is_graph_directed the pairs are de-duplicated as (min,max) and
 * mirrored (enforceBidirectionalization, :218-258); if directed, duplicates are kept.
 * by_source != 0 builds the out-CSR instead (row u = sorted destinations of u), which is what
 * positive sampling walks.
 */
int gigl_graph_from_edges_host(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src,
                               const int32_t* dst, int32_t is_graph_directed, int32_t by_source,
                               gigl_graph** out);
/* Same, from int32 edge arrays already on the device (left untouched). */
int gigl_graph_from_edges_dev(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src_dev,
                              const int32_t* dst_dev, int32_t is_graph_directed, int32_t by_source,
                              gigl_graph** out);
/*
 * Edge-row map of the in-CSR that gigl_graph_from_edges_host builds from the same arguments: edge_rows[j] = index of
 * the input edge record that hydrates CSR slot j (its feature row), what hydrateEdges' join on (_from, _to) looks up
 * (SGSPureSparkV1Task.scala:540-563).  Directed: slot order among equal (dst, src) keys is input order.  Undirected:
 * one record per (least, greatest) pair survives (enforceBidirectionalization :218-258 keeps an arbitrary duplicate;
 * this build keeps the lowest record index) and serves both orientations.  rows_cap = capacity of edge_rows (>= the
 * CSR's edge count); *n_rows = entries written.
 */
int gigl_edge_rows_host(gigl_ctx* ctx, int64_t n_nodes, int64_t n_edges, const int32_t* src, const int32_t* dst,
                        int32_t is_graph_directed, int32_t* edge_rows, int64_t rows_cap, int64_t* n_rows);
int gigl_graph_num_nodes(const gigl_graph* g, int64_t* n_nodes, int64_t* n_edges);
/* Device pointers of the resident CSR (for bindings that want to read it back / reuse it). */
int gigl_graph_device_ptrs(const gigl_graph* g, const int64_t** rowptr_dev, const int32_t** col_dev);
void gigl_graph_destroy(gigl_graph* g);

/* ---- k-hop rooted-neighbourhood index sampling ----------------------------------------- */

/*
 * For every root r and hop h = 1..n_hops, samples min(fanouts[h-1], size) in-neighbours of every
 * vertex on the level-(h-1) frontier with GiGL's deterministic hash permutation:
 *     key_i = XXH64_le32( int32(i + internal_seed + base_seed * (first_call_no + h - 1)) , 42 ),
 *     i = 1..size, the `fanout` smallest (signed key, i) win, emitted in ascending key order,
 * internal_seed = wrapping int32 sum of the path ids root..level h-1.  base_seed = 42 and
 * first_call_no = 1 reproduce SubgraphSamplerTask.samplingSeed / SamplingStrategy._counter.
 *
 * Output = the "padded tree": level h has n_roots * prod(fanouts[0..h-1]) int32 slots in
 * nbr[h-1] (-1 = empty); slot s of level h is child (s % fanouts[h-1]) of slot (s / fanouts[h-1])
 * of level h-1 (level 0 = roots).  cnt[h-1][p] = children written under parent slot p
 * (n_roots * prod(fanouts[0..h-2]) int32 entries).  Sampled edge for a filled slot:
 * src = nbr[h-1][s], dst = its parent (src = _k_hop, dst = _k-1_hop, :615-629).
 * If one parent's sample holds the same vertex m > 1 times (duplicate directed edges), the
 * reference's GROUP BY (_0_hop,_1_hop) yields ONE group over IN(k) repeated m times: it is
 * stored under the FIRST such slot, the other duplicate slots get cnt 0.
 *
 * fanouts[h] in [1, GIGL_MAX_FANOUT], n_hops in [1, GIGL_MAX_HOPS].
 */
/*
 * The sampler keeps a per-graph index of the hash sequence (block-wise smallest keys) so that a
 * hub row costs O(L + size / L) instead of O(size) hash evaluations; results are bit-identical
 * with and without it.  enabled = 0 switches it off (tests compare both paths); default on.
 */
int gigl_graph_set_hash_index(gigl_graph* g, int32_t enabled);

int gigl_sample_khop_host(gigl_graph* g, const int32_t* roots, int64_t n_roots, const int32_t* fanouts,
                          int32_t n_hops, int32_t base_seed, int32_t first_call_no,
                          int32_t* const* nbr /* [n_hops] host */, int32_t* const* cnt /* [n_hops] host */);
/* Device-pointer variant: roots, nbr[h], cnt[h] are device arrays (the pointer tables themselves
 * are host arrays).  Asynchronous on the ctx stream; errors found on the device surface at
 * the next gigl_ctx_sync(). */
int gigl_sample_khop_dev(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                         int32_t n_hops, int32_t base_seed, int32_t first_call_no,
                         int32_t* const* nbr_dev, int32_t* const* cnt_dev);

/*
 * One sampling op of a SamplingOpDAG (subgraph_sampling_strategy.proto:38-58; the per-op expansion of
 * GraphDBSampler.getKHopSubgraphForRootNode, scala_spark35/subgraph_sampler/src/main/scala/libs/sampler/GraphDBSampler.scala:40-148):
 * expands the frontier its parent op produced (or the roots, depth = 1) over ONE edge type, whose CSR `g` holds - by
 * destination for an INCOMING op, by source (by_source build) for an OUTGOING one.  The op sits at `depth` in its chain
 * root -> ... -> op: chain_fanouts[0 .. depth) are the fanouts of its ancestors and its own (last), chain_nbr[0 .. depth-1)
 * the ancestors' padded-tree outputs (the layout of gigl_sample_khop_*).  Output: nbr_out [n_roots * prod chain_fanouts],
 * cnt_out [n_roots * prod chain_fanouts[0 .. depth-1)].  Same permutation as gigl_sample_khop_*: key =
 * XXH64(i + path-id sum + base_seed * call_no), so a linear chain of ops with call_no = 1, 2, ... over one graph equals
 * gigl_sample_khop_* exactly; a tree-shaped DAG gives every op its own call_no (its 1-based position in the DAG).
 * The reference's graph-DB samplers have no reproducible sampling (LocalDbClient.scala:186,204 is `Set.take(n)`), so
 * this is a valid uniform sample, not a bit-level restatement.
 */
int gigl_sample_op_dev(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, int32_t depth, const int32_t* chain_fanouts,
                       const int32_t* const* chain_nbr_dev, int32_t base_seed, int32_t call_no, int32_t* nbr_out_dev,
                       int32_t* cnt_out_dev);
int gigl_sample_op_host(gigl_graph* g, const int32_t* roots, int64_t n_roots, int32_t depth, const int32_t* chain_fanouts,
                        const int32_t* const* chain_nbr, int32_t base_seed, int32_t call_no, int32_t* nbr_out, int32_t* cnt_out);

/*
 * The weighted sampling methods of a SamplingOp (subgraph_sampling_strategy.proto:17-36 `RandomWeighted`, `TopK`), as the
 * reference's Nebula translator words them (scala_spark35/common/src/main/scala/graphdb/nebula/
 * NebulaQueryResponseTranslator.scala:73-105): TopK = "ORDER BY <edgeFeatName> DESC | LIMIT numNodesToSample";
 * RandomWeighted = the same on "<edgeFeatName> * rand()" (its own comment: not true weighted sampling).
 * weights_dev[p] = the op's edge feature of the edge at CSR position p of `g` (gigl_edge_rows_host maps positions to input
 * edge records).  Every frontier slot keeps the numNodesToSample largest scores of its row: score = weight (top-k) or
 * weight * u with u in (0, 1) taken from the op's seeded permutation key of the position, (top 52 bits + 1/2) * 2^-52 - the
 * reference's rand() is unseeded, so RandomWeighted is reproducible here and not there.  Ties go to the lower CSR position
 * (= lower neighbour id; Nebula leaves them unspecified), NaN weights sort last, a frontier node repeated among its siblings
 * is expanded once.  Other arguments and the output layout as gigl_sample_op_dev.
 */
#define GIGL_SAMPLE_UNIFORM 0
#define GIGL_SAMPLE_TOP_K 1
#define GIGL_SAMPLE_RANDOM_WEIGHTED 2
int gigl_sample_op_weighted_dev(gigl_graph* g, const int32_t* roots_dev, int64_t n_roots, int32_t depth, const int32_t* chain_fanouts,
                                const int32_t* const* chain_nbr_dev, const float* weights_dev, int32_t method, int32_t base_seed,
                                int32_t call_no, int32_t* nbr_out_dev, int32_t* cnt_out_dev);

/*
 * Distinct frontier of a SamplingOp (replaces the HashSet[Node] the reference's graph-DB sampler collects the parents'
 * result nodes into before an op's query runs, scala_spark35/subgraph_sampler/src/main/scala/libs/sampler/
 * GraphDBSampler.scala:66-82): out = cur ([n_roots * cur_slots], a parent level in the padded-tree layout) with, per
 * root, every node that already occurs at a lower slot of cur, or anywhere in one of the n_prev (<= 8) earlier lists
 * prev[p] ([n_roots * prev_slots[p]], the op's earlier input instances), replaced by -1.  gigl_sample_op_* expands
 * the copy: each distinct frontier node is then expanded once per op, as the reference does.
 */
int gigl_frontier_distinct_dev(gigl_ctx* ctx, int64_t n_roots, int32_t n_prev, const int32_t* const* prev_dev,
                               const int32_t* prev_slots, const int32_t* cur_dev, int32_t cur_slots, int32_t* out_dev);

/*
 * Positive (out-edge) sampling for node-anchor link prediction: `g_out` is the CSR by SOURCE;
 * for every src u: P(u) = first num_pos of perm(OUT(u), internal_seed = u, call_no) - the NABLP
 * task calls it third, so call_no = 3.  pos: int32[n_srcs * num_pos] (-1 padded), pos_cnt: int32[n_srcs].
 */
int gigl_sample_positives_host(gigl_graph* g_out, const int32_t* srcs, int64_t n_srcs, int32_t num_pos,
                               int32_t base_seed, int32_t call_no, int32_t* pos, int32_t* pos_cnt);

/* ---- aggregate: SAGEConv / GCNConv forward (fp32) --------------------------------------- */

/*
 * edge_index (COO, PyG convention: row 0 = src j, row 1 = dst i, int64, e columns, device) ->
 * CSR by dst with the edges of each row kept in their input order (stable), so the fp32
 * accumulation order equals a sequential index_add_.  rowptr_dev: int64[n+1], col_dev: int32[e].
 * Returns GIGL_E_RANGE at the next sync if an id is outside [0, n).
 */
int gigl_csr_from_coo_dev(gigl_ctx* ctx, int64_t n, int64_t e, const int64_t* src_dev, const int64_t* dst_dev,
                          int64_t* rowptr_dev, int32_t* col_dev);

/*
 * out[i,:] = Wl @ mean_{j in row i} x[j,:] + bl + Wr @ x[i,:]   (mean of an empty row = 0;
 * duplicates counted; optional fused ReLU = the inter-layer activation of BasicGNN).
 * x: [n, F] fp32 row-major, Wl/Wr: [O, F] row-major (PyG lin_l.weight / lin_r.weight), bl: [O] or
 * NULL, out: [n, O].  n_rows_out <= n limits the output to the first n_rows_out rows (rows are
 * still gathered from all n); pass n for the full layer.  All pointers are device pointers.
 */
int gigl_sage_conv_dev(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O,
                       const int64_t* rowptr_dev, const int32_t* col_dev, const float* x_dev,
                       const float* Wl_dev, const float* bl_dev, const float* Wr_dev, float* out_dev,
                       int32_t relu);
/* Host-buffer variant taking the PyG inputs as they are (x, edge_index int64 [2, e]); does
 * H2D, COO->CSR, the layer, D2H. */
int gigl_sage_conv_host(gigl_ctx* ctx, int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* edge_index,
                        const float* x, const float* Wl, const float* bl, const float* Wr, float* out,
                        int32_t relu);
/* Mean aggregation alone (the gather-SpMM without the projection): agg[i,:] = mean_j x[j,:]. */
int gigl_gather_mean_dev(gigl_ctx* ctx, int64_t n_rows_out, int32_t F, const int64_t* rowptr_dev,
                         const int32_t* col_dev, const float* x_dev, float* agg_dev);

/*
 * GCNConv forward: x' = x @ W^T; self loops: every node gets exactly one (existing self loops
 * are collapsed into it); deg_i = 1 + #non-loop in-edges; out_i = sum_j dinv_j dinv_i x'_j + b.
 * CSR rows may contain self loops; they are skipped and replaced by the single implicit loop.
 */
int gigl_gcn_conv_dev(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr_dev,
                      const int32_t* col_dev, const float* x_dev, const float* W_dev, const float* b_dev,
                      float* out_dev, int32_t relu);

/* Host-buffer variant (x, edge_index int64 [2, e] as PyG holds them). */
int gigl_gcn_conv_host(gigl_ctx* ctx, int64_t n, int64_t e, int32_t F, int32_t O, const int64_t* edge_index,
                       const float* x, const float* W, const float* b, float* out, int32_t relu);

/*
 * F.linear on the tensor cores: C[M, N] = A[M, K] @ W[N, K]^T + bias (optional ReLU), fp32 in and
 * out, computed as 3xTF32 on tcgen05 (relative error ~2^-21, inside the 1e-5 parity bound).  This
 * is the weight projection of SAGEConv (lin_l / lin_r) and GCNConv (lin); exposed on its own for
 * the PyTorch binding and the tests.  Row pitches lda / ldw / ldc are in floats.
 */
int gigl_linear_dev(gigl_ctx* ctx, int64_t M, int32_t N, int32_t K, const float* A_dev, int64_t lda, const float* W_dev,
                    int64_t ldw, const float* bias_dev, float* C_dev, int64_t ldc, int32_t relu);

/* ---- aggregate: training forms (forward that keeps its inputs, backward) ------------------- */

/*
 * What `loss.backward()` runs for the layers above when the reference trains
 * (python/gigl/src/common/modeling_task_specs/graphsage_template_modeling_spec.py:299-367,
 * node_classification_modeling_task_spec.py:134-173; torch autograd through torch_geometric SAGEConv / GCNConv).
 *
 * gigl_sage_conv_train_fwd_dev = gigl_sage_conv_dev that also leaves [mean | self] of the n_rows_out output rows in
 * saved_dev: fp32 [n_rows_out, 2 * Fp], Fp = (F + 3) & ~3 (zero padded), which the backward pass re-uses.
 */
int gigl_sage_conv_train_fwd_dev(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O, const int64_t* rowptr_dev,
                                 const int32_t* col_dev, const float* x_dev, const float* Wl_dev, const float* bl_dev,
                                 const float* Wr_dev, float* out_dev, float* saved_dev, int32_t relu);
/*
 * Gradients of gigl_sage_conv_train_fwd_dev.  grad_out: [n_rows_out, O]; out_dev: the forward output (only read when
 * relu != 0, for the mask); rowptr_dev: the forward CSR by dst (row degrees); t_rowptr_dev / t_col_dev: the same edges as
 * a CSR by SOURCE over all n nodes (row j = destinations of j; gigl_csr_from_coo_dev with src / dst swapped) - needed
 * only when grad_x_dev != NULL.  Outputs (any may be NULL to skip, the two weight gradients go together):
 *   grad_x [n, F] = A^T (grad W_l) / deg + grad W_r,  grad_Wl / grad_Wr [O, F],  grad_bl [O].
 * Deterministic: no float atomics anywhere (transposed gather + split-K partials summed in fixed order).
 */
int gigl_sage_conv_bwd_dev(gigl_ctx* ctx, int64_t n, int64_t n_rows_out, int32_t F, int32_t O, const int64_t* rowptr_dev,
                           const int64_t* t_rowptr_dev, const int32_t* t_col_dev, const float* saved_dev, const float* Wl_dev,
                           const float* Wr_dev, const float* out_dev, const float* grad_out_dev, float* grad_x_dev,
                           float* grad_Wl_dev, float* grad_bl_dev, float* grad_Wr_dev, int32_t relu);
/* Gradients of gigl_gcn_conv_dev (x, W as in the forward; out_dev only read when relu != 0). */
int gigl_gcn_conv_bwd_dev(gigl_ctx* ctx, int64_t n, int32_t F, int32_t O, const int64_t* rowptr_dev, const int32_t* col_dev,
                          const int64_t* t_rowptr_dev, const int32_t* t_col_dev, const float* x_dev, const float* W_dev,
                          const float* out_dev, const float* grad_out_dev, float* grad_x_dev, float* grad_W_dev, float* grad_b_dev,
                          int32_t relu);
/*
 * C[M, N] (+)= G[R, M]^T @ A[R, N] on tcgen05 (3xTF32, operands read MN-major straight from their row-major
 * layout, split-K over R with a fixed-order reduction): the weight gradient of F.linear.  Pitches in floats.
 */
int gigl_linear_tn_dev(gigl_ctx* ctx, int64_t R, int32_t M, int32_t N, const float* G_dev, int64_t ldg, const float* A_dev,
                       int64_t lda, float* C_dev, int64_t ldc, int32_t accumulate);

/* ---- resident node features ------------------------------------------------------------- */

/*
 * The graph-wide feature table x[n_nodes, F] fp32 (row v = _node_features of node v, the flattened
 * featureKeys of loadNodeDataframeIntoSparkSql, SGSPureSparkV1Task.scala:90-104).  *_host copies it
 * to HBM (owned by the graph); *_dev wraps a device array the caller keeps alive.
 */
int gigl_graph_set_features_host(gigl_graph* g, const float* x, int32_t F);
int gigl_graph_set_features_dev(gigl_graph* g, const float* x_dev, int32_t F);
/* Same with a row pitch (floats) >= F.  A pitch that is a multiple of 32 floats keeps every row on whole 128-byte lines,
 * which is what a feature table sharded over NVLink wants (gigl_shared_table_*): remote rows then cross the link as full
 * lines (measured at F = 100 vs 128 on two B200s: 369 vs 742 GB/s of remote row traffic). */
int gigl_graph_set_features_pitched_dev(gigl_graph* g, const float* x_dev, int32_t F, int64_t row_pitch);
int gigl_graph_features_dev(const gigl_graph* g, const float** x_dev, int32_t* F);

/* ---- node features sharded over the GPUs of one NVSwitch box ------------------------------- */

/*
 * the feature table is sharded by contiguous node-id range, shard k on GPU k, and stitched
 * into ONE flat virtual array on every GPU with the CUDA virtual memory API (cuMemCreate / cuMemMap): row v lives at
 * base + v * F floats everywhere, a remote row is an ordinary load routed over NVLink.  The result is handed to
 * gigl_graph_set_features_dev / gigl_batch_sage_forward_dev like any device array.  One process per GPU:
 *   1. every rank: gigl_shared_table_create(... my_shard = rank ...) -> a POSIX file descriptor of its shard;
 *   2. the ranks exchange the descriptors (Unix socket SCM_RIGHTS; gigl_b200/sharding.py);
 *   3. every rank: gigl_shared_table_attach(shard k, fd of rank k) for all k != rank.
 * rows_per_shard must be a multiple of gigl_shared_table_row_granule(F) (physical allocations are mapped at the
 * driver's allocation granularity, typically 2 MiB).
 */
typedef struct gigl_shared_table gigl_shared_table;
int gigl_shared_table_row_granule(gigl_ctx* ctx, int32_t F, int64_t* rows);
int gigl_shared_table_create(gigl_ctx* ctx, int32_t n_shards, int32_t my_shard, int64_t rows_per_shard, int32_t F,
                             gigl_shared_table** out, int32_t* export_fd);
int gigl_shared_table_attach(gigl_shared_table* t, int32_t shard, int32_t fd);
/* base of the flat table [n_shards * rows_per_shard, F], the start of this rank's own shard, total rows */
int gigl_shared_table_ptrs(const gigl_shared_table* t, float** base_dev, float** my_shard_dev, int64_t* total_rows);
void gigl_shared_table_destroy(gigl_shared_table* t);

/* ---- model: GraphSAGE weights resident on the device ------------------------------------ */

typedef struct gigl_sage_model gigl_sage_model;
/*
 * torch_geometric.nn.GraphSAGE(in, hidden, num_layers, out) as the reference builds it
 * (graphsage_template_modeling_spec.py:143-148; homogeneous.py:171-202): layer l maps dims[l] ->
 * dims[l+1]; Wl[l] = convs.{l}.lin_l.weight [dims[l+1], dims[l]], bl[l] = convs.{l}.lin_l.bias
 * (or NULL), Wr[l] = convs.{l}.lin_r.weight.  ReLU between layers, none after the last.
 * The weights are copied (host or device source) and re-laid out as [Wl | Wr] per layer.
 */
int gigl_sage_model_create_host(gigl_ctx* ctx, int32_t n_layers, const int32_t* dims, const float* const* Wl,
                                const float* const* bl, const float* const* Wr, gigl_sage_model** out);
int gigl_sage_model_create_dev(gigl_ctx* ctx, int32_t n_layers, const int32_t* dims, const float* const* Wl_dev,
                               const float* const* bl_dev, const float* const* Wr_dev, gigl_sage_model** out);
void gigl_sage_model_destroy(gigl_sage_model* m);

/* ---- batch: B sampled neighbourhoods -> one coalesced graph -> root embeddings ----------- */

typedef struct gigl_batch gigl_batch;
/*
 * A reusable collation workspace for graphs of n_graph_nodes vertices (dense per-vertex maps in
 * HBM: 12 bytes x n_graph_nodes).  Replaces the per-batch Python graph building of
 * python/gigl/src/common/graph_builder/pyg_graph_builder.py:20-69 and the collate functions of
 * python/gigl/src/training/v1/lib/data_loaders/ (see gigl_b200/csrc/batch_collate.cu).
 */
int gigl_batch_create(gigl_ctx* ctx, int64_t n_graph_nodes, gigl_batch** out);
void gigl_batch_destroy(gigl_batch* b);
/*
 * Coalesces the padded-tree sample of gigl_sample_khop_dev (same roots / fanouts / nbr tables, all
 * on the device) into the batch graph: nodes de-duplicated by id, edges de-duplicated by
 * (src, dst).  n_layers = the depth of the model that will run on it: local ids are ordered so
 * that the first level_sizes[j] nodes are exactly those whose layer-(n_layers - j) output the root
 * embeddings depend on (level_sizes[0] = n_roots).  Synchronises the stream once to return the
 * sizes.  level_sizes: host int64[n_layers]; n_edges: unique edges of the batch graph.
 */
int gigl_batch_collate_dev(gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                           int32_t n_hops, const int32_t* const* nbr_dev, int32_t n_layers, int64_t* level_sizes,
                           int64_t* n_edges);
/*
 * The batch graph as the reference's PygGraphBuilder would hand it to the model.
 * gigl_batch_finalize_nodes gives every remaining batch node a local id (the levels above only
 * cover what the root outputs need) and returns the sizes; gigl_batch_export_dev then writes
 * node_ids int32[n_nodes] (local id -> global id; roots first) and edge_index int64[2, n_edges]
 * in LOCAL ids (row 0 = src, row 1 = dst; sorted by (dst, src) global id), both device arrays.
 */
int gigl_batch_finalize_nodes(gigl_batch* b, int64_t* n_nodes, int64_t* n_edges);
int gigl_batch_export_dev(gigl_batch* b, int32_t* node_ids_dev, int64_t* edge_index_dev);
/*
 * model(x, edge_index)[root rows] on the collated batch: out_dev [n_roots, dims[n_layers]] fp32.
 * x_dev is the graph-wide feature table indexed by GLOBAL node id (row stride ldx floats).
 */
int gigl_batch_sage_forward_dev(gigl_batch* b, const gigl_sage_model* m, const float* x_dev, int64_t ldx,
                                float* out_dev);
/*
 * The remote-neighbour feature halo of a sharded feature table (gigl_shared_table_*).  Off
 * (default): the layer-1 gather loads every source row straight from x_dev - one row per unique batch EDGE, remote rows
 * over NVLink.  On: gigl_batch_sage_forward_dev first copies the row of every unique batch NODE into a per-batch table in
 * local HBM (one row per node crosses NVLink; this is the ids -> rows exchange an all_to_all halo performs, done with
 * peer loads) and layer 1 gathers from that copy through the local-id map.  Same embeddings bit for bit; worth it when
 * most rows are remote (a batch has several times more unique edges than unique nodes), a loss on a local table.
 */
int gigl_batch_set_halo_staging(gigl_batch* b, int32_t enabled);
/*
 * Early staging: registers the feature table the staged halo copies from (x_dev [n_graph_nodes, F], row pitch ldx floats;
 * the table gigl_batch_sage_forward_dev will be called with; must outlive the batch; NULL unregisters).  With a registered
 * table gigl_batch_collate_dev forks the halo onto a side stream as soon as it is called: the rows a batch needs are the
 * roots and every filled slot of the sampled tree, known BEFORE the collation, so a stage slot is claimed per distinct
 * vertex and the rows are copied (NVLink for the remote ones) WHILE the collation's kernels run on the context's stream;
 * the forward joins the copy before layer 1 and gathers through the stage-slot map.  Same embeddings bit for bit (the same
 * rows summed in the same order); the exposed cost of the halo drops by the duration of the collation.
 */
int gigl_batch_set_halo_table_dev(gigl_batch* b, const float* x_dev, int32_t F, int64_t ldx);
/*
 * gigl_sample_khop_dev that starts the halo even earlier: `b` is the batch workspace the sample will be collated into next,
 * with a table registered by gigl_batch_set_halo_table_dev that is also the graph's feature table.  The hops are sampled one
 * launch sequence at a time and after each one the rows of the level just sampled (first the roots) are claimed and copied
 * on the batch's side stream: hop h's rows cross NVLink under the sampling of hop h + 1, the last hop's under the collation.
 * Without a registered table it is gigl_sample_khop_dev.  Same index sets, same embeddings.
 */
int gigl_sample_khop_staged_dev(gigl_graph* g, gigl_batch* b, const int32_t* roots_dev, int64_t n_roots, const int32_t* fanouts,
                                int32_t n_hops, int32_t base_seed, int32_t first_call_no, int32_t* const* nbr_dev,
                                int32_t* const* cnt_dev);
/*
 * Hot rows of the staged halo: hot_dev [n_hot, ld] holds a LOCAL copy of the feature rows of the vertices batches meet
 * most often (the caller picks them - gigl_b200.sharding.hot_rows takes the highest-degree vertices - and fills the copy
 * once, from the sharded table), hot_slot_dev [n_graph_nodes] maps a vertex to its row there (-1 = not replicated).
 * The staging kernel then reads those rows from local HBM and only the cold tail of a batch crosses NVLink; the
 * embeddings are unchanged (the same bytes from another address).  NULL / NULL switches it off.  Both arrays stay owned
 * by the caller and must outlive the batch's forward calls.  This is the replicate-the-hubs / shard-the-tail residency
 * between "replicated" and "sharded" of SURVEY.md section 8(e): per-GPU memory = 1/N of the table + the hot fraction.
 */
int gigl_batch_set_hot_rows_dev(gigl_batch* b, const int32_t* hot_slot_dev, const float* hot_dev, int32_t F, int64_t ld);

/*
 * One call from host buffers: roots (host) -> k-hop sample -> collate -> GraphSAGE forward ->
 * root embeddings (host).  The graph must hold features (gigl_graph_set_features_*).  If nbr / cnt
 * are non-NULL the padded-tree index sets are returned too (layout of gigl_sample_khop_host).
 * This is the entry point a JNI / ctypes binding of the reference's sampler + inferencer pair
 * calls per batch (KHopSamplerService.getKHopSubgraphForRootNodes + BaseInferencer.infer_batch,
 * python/gigl/src/inference/v1/lib/base_inferencer.py:23-57).
 */
int gigl_infer_khop_sage_host(gigl_graph* g, gigl_batch* b, const gigl_sage_model* m, const int32_t* roots,
                              int64_t n_roots, const int32_t* fanouts, int32_t n_hops, int32_t base_seed,
                              int32_t first_call_no, float* out, int32_t* const* nbr, int32_t* const* cnt);

/*
 * Same call, index sets returned PACKED (csrc/tree_pack.cu): cnt_u8[h][p] = number of children under parent slot p of
 * hop h + 1 as one byte (fanout <= 128; n_roots * prod(fanouts[0..h-1]) entries), packed = the filled slots only - hop
 * after hop, parent slots in order, each parent's children in the order of the padded layout (a parent's filled slots
 * are its first cnt slots), *n_packed entries in all (<= packed_cap, else GIGL_E_INVALID).  The padded layout of
 * gigl_sample_khop_host is recovered by writing each parent's cnt children back at slot p * fanout (gigl_b200.engine.
 * unpack_tree).  Same sampled edges, ~2/3 of the bytes on the products-like graph: with 8 ranks on one host the
 * device-to-host copies, not the GPUs, bound the end-to-end rate.  This is the un-padded Seq a
 * KHopSamplerService.getKHopSubgraphForRootNodes returns (KHopSamplerService.scala:17-20).
 */
int gigl_infer_khop_sage_packed_host(gigl_graph* g, gigl_batch* b, const gigl_sage_model* m, const int32_t* roots, int64_t n_roots,
                                     const int32_t* fanouts, int32_t n_hops, int32_t base_seed, int32_t first_call_no, float* out,
                                     uint8_t* const* cnt_u8 /* [n_hops] host */, int32_t* packed, int64_t packed_cap, int64_t* n_packed);

/*
 * The packed form again with the ids as a BIT STREAM: every filled slot takes *id_bits = ceil(log2(n_nodes)) bits (22 instead
 * of 32 on a 2.4 M-node graph), entry e in bits [e * id_bits, (e + 1) * id_bits) of the little-endian 32-bit words; counts
 * and order as in gigl_infer_khop_sage_packed_host.  *n_packed = entries (filled slots); ceil(n_packed * id_bits / 32) words
 * are written (<= words_cap, else GIGL_E_INVALID).  gigl_unpack_bits_host (pure host code) turns the stream back into int32
 * ids.  Why: on an 8-GPU host the device-to-host copies bound the end-to-end rate (about 100 GB/s for the whole box measured,
 * 12 GB/s per GPU), so bytes returned are throughput.
 */
int gigl_infer_khop_sage_bitpacked_host(gigl_graph* g, gigl_batch* b, const gigl_sage_model* m, const int32_t* roots, int64_t n_roots,
                                        const int32_t* fanouts, int32_t n_hops, int32_t base_seed, int32_t first_call_no, float* out,
                                        uint8_t* const* cnt_u8, uint32_t* words, int64_t words_cap, int64_t* n_packed, int32_t* id_bits);
int gigl_unpack_bits_host(const uint32_t* words, int64_t n, int32_t bits, int32_t* out);

/* ---- the sampler's file contract: TFRecord + tf.Example + sample protos (host code) ----------- */

/* masked crc32c of TFRecord framing: rotr15(crc32c(data)) + 0xA282EAD8 */
uint32_t gigl_crc32c_masked(const void* data, int64_t n);
/* frees buffers returned by gigl_encode_samples_host */
void gigl_free_host(void* p);
/*
 * Hydrates the padded-tree index sets of gigl_sample_khop_host into serialized sample protos, one
 * per root, in root order:
 *   kind 0: RootedNodeNeighborhood { root_node, neighborhood { nodes, edges } }
 *   kind 1: SupervisedNodeClassificationSample { ..., root_node_labels [{label_type, label}] } for the
 *           roots that carry a label (labels[node] != INT32_MIN), as the reference's inner join does
 * (proto/snapchat/research/gbml/training_samples_schema.proto:16-31; reference: SGSPureSparkV1Task.scala:496-820,
 * 1019-1040, SupervisedNodeClassificationTask.scala:166-236).  nodes = array_distinct(hop nodes ++ root), every node
 * with its feature row x[node, :] (x may be NULL with F = 0) and condensed_node_type (>= 0, or -1 to leave it
 * unset); edges = one per filled tree slot, src = hop-k node, dst = its parent (:615-629), with
 * condensed_edge_type.  tfrecord_framing != 0 wraps every message as a TFRecord (what
 * TFRecordIO.writeDatasetToTfrecord emits, TFRecordIO.scala:53-69).  *out is malloc'd (gigl_free_host);
 * record_offsets (optional, n_roots + 1) gives each root's byte range (empty for skipped roots).
 */
int gigl_encode_samples_host(int32_t kind, int64_t n_roots, const int32_t* roots, const int32_t* fanouts, int32_t n_hops,
                             const int32_t* const* nbr, const float* x, int32_t F, int32_t condensed_node_type,
                             int32_t condensed_edge_type, const int32_t* labels, const char* label_type, int32_t tfrecord_framing,
                             uint8_t** out, int64_t* out_bytes, int64_t* record_offsets);
/*
 * Extended form: edge hydration and the link-prediction sample type.
 *   rowptr / col (host copies of the in-CSR, optional): every sampled pair (src -> dst) becomes one Edge per matching
 *     edge record, which is hydrateEdges' INNER JOIN on (_from, _to) (SGSPureSparkV1Task.scala:540-563): one for
 *     undirected graphs, one per duplicate record for directed graphs that carry duplicates.  With edge_feat
 *     [n_records, Fe] the Edge carries feature_values = edge_feat[edge_rows[slot]] (edge_rows from gigl_edge_rows_host;
 *     NULL = the CSR slot index itself).
 *   kind 2: NodeAnchorBasedLinkPredictionSample { root_node, pos_edges, neighborhood } for the first n_emit roots (the
 *     anchors); roots[n_emit..n_roots) only supply the trees of positives that are not anchors of this call.
 *     pos [n_emit * num_pos] = sampled positive destinations (gigl_sample_positives_host; -1 = none), pos_tree = index
 *     into roots of each positive's tree (-1 = none).  neighborhood = array_distinct(the anchor's ++ every positive's)
 *     over nodes and over hydrated edges (NodeAnchorBasedLinkPredictionTask.scala:186-209,
 *     NodeAnchorBasedLinkPredictionBaseTask.scala:106-198); pos_edges = (anchor -> positive), hydrated (:280-334);
 *     hard_neg_edges / neg_edges stay empty (:388-406).  Anchors without a positive emit nothing.
 * For kinds 0 / 1 pass n_emit = n_roots.  record_offsets has n_emit + 1 entries.
 */
int gigl_encode_samples_ex_host(int32_t kind, int64_t n_roots, int64_t n_emit, const int32_t* roots, const int32_t* fanouts,
                                int32_t n_hops, const int32_t* const* nbr, const float* x, int32_t F, int32_t condensed_node_type,
                                int32_t condensed_edge_type, const int64_t* rowptr, const int32_t* col, const int32_t* edge_rows,
                                const float* edge_feat, int32_t Fe, const int32_t* labels, const char* label_type, int32_t num_pos,
                                const int32_t* pos, const int64_t* pos_tree, int32_t tfrecord_framing, uint8_t** out,
                                int64_t* out_bytes, int64_t* record_offsets);
/*
 * One hydrated edge table as the sample encoder joins against it: the in-CSR by destination (host copies; rows
 * ascending, duplicates kept), the input record behind every CSR slot (gigl_edge_rows_host; NULL = the slot index) and
 * the records' feature rows [n_records, n_feat].  rowptr == NULL: every sampled pair matches exactly one feature-less
 * record.
 */
typedef struct gigl_edge_table {
    const int64_t* rowptr;
    const int32_t* col;
    const int32_t* edge_rows;
    const float* feat;
    int32_t n_feat;
} gigl_edge_table;
/*
 * NodeAnchorBasedLinkPredictionSample with user-defined labels
 * (UserDefinedLabelsNodeAnchorBasedLinkPredictionTask.scala:54-581): as kind 2 of gigl_encode_samples_ex_host, but the
 * positives were sampled from - and pos_edges are hydrated against - pos_edges (NULL = the main table, i.e. sampled
 * from the graph's own out-edges), and up to num_neg hard negatives per anchor from neg_edges fill hard_neg_edges
 * (neg / neg_tree laid out like pos / pos_tree; num_neg = 0: none).  An anchor needs a positive; negatives are
 * optional (LEFT JOIN, :405-430).  The neighbourhood is array_distinct(anchor's ++ positives' ++ negatives').
 */
int gigl_encode_link_samples_host(int64_t n_roots, int64_t n_emit, const int32_t* roots, const int32_t* fanouts, int32_t n_hops,
                                  const int32_t* const* nbr, const float* x, int32_t F, int32_t condensed_node_type,
                                  int32_t condensed_edge_type, const gigl_edge_table* main_edges, const gigl_edge_table* pos_edges,
                                  const gigl_edge_table* neg_edges, int32_t num_pos, const int32_t* pos, const int64_t* pos_tree,
                                  int32_t num_neg, const int32_t* neg, const int64_t* neg_tree, int32_t tfrecord_framing, uint8_t** out,
                                  int64_t* out_bytes, int64_t* record_offsets);
/*
 * Typed (heterogeneous) RootedNodeNeighborhoods from the ops of a SamplingOp DAG (gigl_sample_op_*): per root the union
 * of the ops' edge and node SETS plus the root (GraphDBSampler.getKHopSubgraphForRootNode,
 * scala_spark35/subgraph_sampler/src/main/scala/libs/sampler/GraphDBSampler.scala:129-148), every node carrying its
 * condensed node type and that type's feature row (SGSTask.hydrateRnn, .../libs/utils/SGSTask.scala:200-337), every
 * edge its condensed edge type (edge features are not hydrated here).  ops are in topological order (parent < own index).
 */
typedef struct gigl_dag_op {
    int32_t parent;              /* index of the input op, -1 = the op expands the root */
    int32_t fanout;
    int32_t condensed_edge_type;
    int32_t result_node_type;    /* condensed node type of the sampled nodes */
    int32_t outgoing;            /* 0 = INCOMING: Edge(sampled -> frontier node); 1 = OUTGOING: Edge(frontier node -> sampled) */
    const int32_t* nbr;          /* the op's padded-tree output (host), n_roots * prod(fanouts along its chain) */
} gigl_dag_op;
typedef struct gigl_node_table {
    const float* x;              /* [n_nodes_of_type, n_feat] (host), NULL with n_feat = 0 */
    int32_t n_feat;
} gigl_node_table;
int gigl_encode_dag_samples_host(int64_t n_roots, const int32_t* roots, int32_t root_node_type, int32_t n_ops, const gigl_dag_op* ops,
                                 int32_t n_node_types, const gigl_node_table* node_tables /* by condensed node type */,
                                 int32_t tfrecord_framing, uint8_t** out, int64_t* out_bytes, int64_t* record_offsets);
/*
 * The typed task's two outputs with edge hydration (GraphDBNodeAnchorBasedLinkPredictionTask.run,
 * scala_spark35/subgraph_sampler/src/main/scala/libs/task/graphdb/GraphDBNodeAnchorBasedLinkPredictionTask.scala:100-495).
 * A gigl_dag_tree is one batch of roots of ONE node type with the padded-tree outputs of that type's SamplingOp DAG.  An op
 * with several input ops appears once per input (one gigl_dag_op per (op, parent) pair): the union of those entries is
 * the op's result set, as GraphDBSampler.scala:66-82 expands the union of its parents' result nodes.
 *   kind 0: RootedNodeNeighborhood per anchor root = gigl_encode_dag_samples_host plus edge hydration.
 *   kind 2: NodeAnchorBasedLinkPredictionSample per anchor root: pos_edges = the distinct (anchor -> positive) edges of
 *     pos [n_roots * num_pos] (-1 = none; sampled by an OUTGOING op over the supervision edge type,
 *     GraphDBSampler.samplePositiveEdgeNeighborhoods :175-215) with condensed type pos_condensed_edge_type;
 *     neighborhood = the anchor's DAG merged BY KEY with the DAG of every positive (pos_tree = index of the positive in
 *     targets->roots, -1 = the positive node alone) - mergeGraphs (GraphPbWrappers.scala:43-68) keeps one Edge per
 *     (src, dst, type).  Anchors without a positive emit nothing unless include_isolated != 0
 *     (sharedConfig.shouldIncludeIsolatedNodesInTraining, the LEFT JOIN at :384-402).
 * edge_tables[t] (optional) = the records of condensed edge type t as an in-CSR by destination (+ feature rows):
 * hydrate_flags bit 0 joins the neighbourhood's edges against it (SGSTask.hydrateRnn's LEFT JOIN on (_from, _to, type):
 * kind 0 emits one Edge per matching record, kind 2 the first), bit 1 the pos_edges (:349-377); an edge without a record or
 * a table stays feature-less.  record_offsets has anchors->n_roots + 1 entries.
 */
typedef struct gigl_dag_tree {
    int64_t n_roots;
    const int32_t* roots;
    int32_t root_node_type;      /* condensed node type of the roots */
    int32_t n_ops;
    const gigl_dag_op* ops;      /* topological order */
} gigl_dag_tree;
int gigl_encode_typed_samples_host(int32_t kind, const gigl_dag_tree* anchors, const gigl_dag_tree* targets, int32_t num_pos,
                                   const int32_t* pos, const int64_t* pos_tree, int32_t pos_condensed_edge_type, int32_t include_isolated,
                                   int32_t hydrate_flags, int32_t n_node_types, const gigl_node_table* node_tables, int32_t n_edge_types,
                                   const gigl_edge_table* edge_tables /* by condensed edge type, or NULL */, int32_t tfrecord_framing,
                                   uint8_t** out, int64_t* out_bytes, int64_t* record_offsets);
/*
 * Splits a TFRecord byte stream into records (payload offsets / lengths, arrays of capacity max_records; pass NULL
 * arrays to only count).  verify != 0 checks both masked crc32c fields.  Returns the record count or GIGL_E_*.
 */
int64_t gigl_tfrecord_index_host(const uint8_t* data, int64_t n_bytes, int32_t verify, int64_t* offsets, int64_t* lengths,
                                 int64_t max_records);
/*
 * TaskOutputValidator.validateRootedNodeNeighborhoodSamples / validateMainSamples (scala/subgraph_sampler/src/main/scala/
 * libs/task/TaskOutputValidator.scala:29-108) on the bytes about to be written: every sample of a TFRecord stream is parsed
 * back and both endpoints of every neighbourhood edge (kind 2: also of every pos / neg / hard-neg edge) must be among the
 * neighbourhood's nodes, compared as (node id, condensed node type) with the node types the edge's condensed edge type
 * implies (edge_src_type / edge_dst_type per condensed edge type; NULL = homogeneous, type 0); a sample without a
 * neighbourhood fails.  kind: 0 RootedNodeNeighborhood, 1 SupervisedNodeClassificationSample, 2 NodeAnchorBasedLink-
 * PredictionSample.  GIGL_OK, or GIGL_E_INVALID with *bad_record = the first offending record and *reason = 1 malformed
 * bytes, 2 neighbourhood missing, 3 endpoint outside the neighbourhood nodes (the reference throws RuntimeException).
 */
int gigl_validate_samples_host(const uint8_t* data, int64_t n_bytes, int32_t kind, int32_t n_edge_types, const int32_t* edge_src_type,
                               const int32_t* edge_dst_type, int64_t* n_records_out, int64_t* bad_record, int32_t* reason);
/*
 * Decodes feature `name` of every tf.Example record into a dense column of `width` values per record: dtype 0 = int64
 * (out_i64), dtype 1 = float (out_f32; an Int64List is cast, as `cast(col as array<float>)` does at
 * SGSPureSparkV1Task.scala:90-104).  GIGL_E_RANGE if a record lacks the feature.
 */
int gigl_examples_column_host(const uint8_t* data, int64_t n_records, const int64_t* offsets, const int64_t* lengths,
                              const char* name, int32_t dtype, int32_t width, int64_t* out_i64, float* out_f32);

#ifdef __cplusplus
}
#endif
#endif /* GIGL_B200_H_ */
