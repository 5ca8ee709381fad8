// sample_io.cu - host side of the sampler's file contract (no device code in this file):
//   * TFRecord framing (u64 length, masked crc32c of the length, payload, masked crc32c of the payload) as
//     written by the reference through the Spark TFRecord connector with recordType=ByteArray
//     (scala/common/src/main/scala/utils/TFRecordIO.scala:53-69) and read with tf.data.TFRecordDataset
//     (python/gigl/src/training/v1/lib/data_loaders/tf_records_iterable_dataset.py:78);
//   * tf.Example decoding of the preprocessed node / edge tables
//     (loadNodeDataframeIntoSparkSql / loadEdgeDataframeIntoSparkSql, SGSPureSparkV1Task.scala:52-311);
//   * hydration + protobuf encoding of the sampled index sets into RootedNodeNeighborhood /
//     SupervisedNodeClassificationSample messages
//     (proto/snapchat/research/gbml/training_samples_schema.proto:16-31, graph_schema.proto:5-31;
//     reference: createKthHydratedNeighborhood / createSubgraph / castToRootedNodeNeighborhoodProtoSchema,
//     SGSPureSparkV1Task.scala:496-820, 1019-1040, and SupervisedNodeClassificationTask.scala:166-236, 320-337).
// Output volume (about (1 + f1 + f1*f2) * F * 4 bytes per root) makes this a host-bandwidth job: the
// encoder is two OpenMP passes over the roots (sizes, then bytes) into one contiguous buffer.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__linux__)
#include <sys/mman.h>
#endif

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/gigl_b200.h"

namespace {

// ---- crc32c (Castagnoli), slicing-by-8 tables + SSE4.2 when the CPU has it ---------------------
uint32_t g_crc_tab[8][256];
bool g_crc_init = false;

void crc_init() {
    if (g_crc_init) return;
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
        g_crc_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t) g_crc_tab[t][i] = (g_crc_tab[t - 1][i] >> 8) ^ g_crc_tab[0][g_crc_tab[t - 1][i] & 0xFF];
    g_crc_init = true;
}

uint32_t crc32c_sw(uint32_t crc, const uint8_t* p, size_t n) {
    crc = ~crc;
    while (n >= 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        v ^= crc;
        crc = g_crc_tab[7][v & 0xFF] ^ g_crc_tab[6][(v >> 8) & 0xFF] ^ g_crc_tab[5][(v >> 16) & 0xFF] ^ g_crc_tab[4][(v >> 24) & 0xFF] ^
              g_crc_tab[3][(v >> 32) & 0xFF] ^ g_crc_tab[2][(v >> 40) & 0xFF] ^ g_crc_tab[1][(v >> 48) & 0xFF] ^ g_crc_tab[0][v >> 56];
        p += 8;
        n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ g_crc_tab[0][(crc ^ *p++) & 0xFF];
    return ~crc;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t crc32c_hw(uint32_t crc, const uint8_t* p, size_t n) {
    uint64_t c = (uint32_t)~crc;
    while (n >= 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        c = __builtin_ia32_crc32di(c, v);
        p += 8;
        n -= 8;
    }
    uint32_t c32 = (uint32_t)c;
    while (n--) c32 = __builtin_ia32_crc32qi(c32, *p++);
    return ~c32;
}
#endif

uint32_t crc32c(const uint8_t* p, size_t n) {
#if defined(__x86_64__)
    static const bool hw = __builtin_cpu_supports("sse4.2");
    if (hw) return crc32c_hw(0, p, n);
#endif
    crc_init();
    return crc32c_sw(0, p, n);
}

inline uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xA282EAD8u; }

// Output buffer of an encoder call (released with gigl_free_host = free).  Gigabyte-sized and written exactly once, so
// first-touch page faults are a real share of the call: large buffers are 2 MiB-aligned and offered to the kernel as
// transparent huge pages (512x fewer faults where THP is in `madvise` or `always` mode; harmless elsewhere).
uint8_t* alloc_out(size_t bytes) {
    if (bytes == 0) bytes = 1;
#if defined(__linux__)
    constexpr size_t kHuge = (size_t)2 << 20;
    if (bytes >= 4 * kHuge) {
        void* p = nullptr;
        if (posix_memalign(&p, kHuge, (bytes + kHuge - 1) & ~(kHuge - 1)) == 0) {
            madvise(p, (bytes + kHuge - 1) & ~(kHuge - 1), MADV_HUGEPAGE);
            return (uint8_t*)p;
        }
    }
#endif
    return (uint8_t*)malloc(bytes);
}

// ---- protobuf wire helpers ----------------------------------------------------------------------
inline int varint_size(uint64_t v) {
    int n = 1;
    while (v >= 0x80) {
        v >>= 7;
        ++n;
    }
    return n;
}
inline uint8_t* put_varint(uint8_t* p, uint64_t v) {
    while (v >= 0x80) {
        *p++ = (uint8_t)(v | 0x80);
        v >>= 7;
    }
    *p++ = (uint8_t)v;
    return p;
}
inline bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t& v) {
    v = 0;
    for (int shift = 0; shift < 64 && p < end; shift += 7) {
        const uint8_t b = *p++;
        v |= (uint64_t)(b & 0x7F) << shift;
        if (!(b & 0x80)) return true;
    }
    return false;
}

// Node { uint32 node_id = 1; optional uint32 condensed_node_type = 2; repeated float feature_values = 3 [packed]; }
inline size_t node_size(uint32_t id, int32_t type, int F) {
    size_t n = 0;
    if (id != 0) n += 1 + varint_size(id);  // proto3 implicit presence: 0 is not written
    if (type >= 0) n += 1 + varint_size((uint32_t)type);
    if (F > 0) n += 1 + varint_size((uint64_t)F * 4) + (size_t)F * 4;
    return n;
}
inline uint8_t* put_node(uint8_t* p, uint32_t id, int32_t type, const float* feat, int F) {
    if (id != 0) {
        *p++ = 0x08;
        p = put_varint(p, id);
    }
    if (type >= 0) {
        *p++ = 0x10;
        p = put_varint(p, (uint32_t)type);
    }
    if (F > 0) {
        *p++ = 0x1A;
        p = put_varint(p, (uint64_t)F * 4);
        memcpy(p, feat, (size_t)F * 4);  // little-endian host
        p += (size_t)F * 4;
    }
    return p;
}
// Edge { uint32 src_node_id = 1; uint32 dst_node_id = 2; optional uint32 condensed_edge_type = 3; repeated float feature_values = 4 [packed]; }
inline size_t edge_size(uint32_t src, uint32_t dst, int32_t type, int Fe) {
    size_t n = 0;
    if (src != 0) n += 1 + varint_size(src);
    if (dst != 0) n += 1 + varint_size(dst);
    if (type >= 0) n += 1 + varint_size((uint32_t)type);
    if (Fe > 0) n += 1 + varint_size((uint64_t)Fe * 4) + (size_t)Fe * 4;
    return n;
}
inline uint8_t* put_edge(uint8_t* p, uint32_t src, uint32_t dst, int32_t type, const float* feat, int Fe) {
    if (src != 0) {
        *p++ = 0x08;
        p = put_varint(p, src);
    }
    if (dst != 0) {
        *p++ = 0x10;
        p = put_varint(p, dst);
    }
    if (type >= 0) {
        *p++ = 0x18;
        p = put_varint(p, (uint32_t)type);
    }
    if (Fe > 0) {
        *p++ = 0x22;
        p = put_varint(p, (uint64_t)Fe * 4);
        memcpy(p, feat, (size_t)Fe * 4);
        p += (size_t)Fe * 4;
    }
    return p;
}

// Host-side tables the hydration joins read (all optional except the trees).
struct ETab {  // one hydrated edge table (main / user-defined positive / user-defined negative)
    const int64_t* rowptr;     // in-CSR by destination, rows ascending (nullptr: one feature-less Edge per sampled pair)
    const int32_t* col;
    const int32_t* edge_rows;  // CSR slot -> row of ef (nullptr: the slot index itself)
    const float* ef;
    int Fe;
};

struct Tables {
    const int32_t* roots;
    const int32_t* fanouts;
    int n_hops;
    const int32_t* const* nbr;
    const float* x;
    int F;
    int32_t ntype, etype;
    ETab tab[3];  // 0 = main edges, 1 = positive label edges, 2 = negative label edges
};

struct EdgeRef {
    uint32_t src, dst;
    int64_t row;  // row of the table's feature matrix, -1 = none
    int tag;      // which ETab hydrates it
};

struct RootPlan {
    std::vector<uint32_t> nodes;     // distinct, first-seen order
    std::vector<EdgeRef> edges;      // src = hop-k node, dst = hop-(k-1) node
    std::vector<EdgeRef> pos_edges;  // NodeAnchorBasedLinkPredictionSample.pos_edges
    std::vector<EdgeRef> neg_edges;  // NodeAnchorBasedLinkPredictionSample.hard_neg_edges
    std::vector<std::pair<uint32_t, uint32_t>> tmp;
};

// hydrateEdges / hydrateTaskBasedEdges: the sampled pair (src -> dst) INNER JOINs the hydrated edge table on
// (_from, _to) (SGSPureSparkV1Task.scala:540-563, NodeAnchorBasedLinkPredictionBaseTask.scala:280-334): one Edge per
// matching record - exactly one for undirected graphs, one per duplicate record for directed tables with duplicates.
void join_edge(const Tables& t, int tag, uint32_t src, uint32_t dst, std::vector<EdgeRef>& out) {
    const ETab& e = t.tab[tag];
    if (!e.rowptr) {
        out.push_back({src, dst, -1, tag});
        return;
    }
    const int32_t* b = e.col + e.rowptr[dst];
    const int32_t* en = e.col + e.rowptr[dst + 1];
    auto r = std::equal_range(b, en, (int32_t)src);
    for (const int32_t* q = r.first; q != r.second; ++q) {
        const int64_t slot = q - e.col;
        out.push_back({src, dst, e.Fe > 0 ? (e.edge_rows ? (int64_t)e.edge_rows[slot] : slot) : -1, tag});
    }
}

// Walks the padded tree of roots[r] (layout: include/gigl_b200.h), appending its edges and (non-distinct) nodes.
void walk_tree(const Tables& t, int64_t r, RootPlan& out) {
    int64_t width_prev = 1;
    for (int h = 0; h < t.n_hops; ++h) {
        const int f = t.fanouts[h];
        const int32_t* cur = t.nbr[h];
        for (int64_t ps = 0; ps < width_prev; ++ps) {
            const int64_t pslot = r * width_prev + ps;
            const int32_t parent = (h == 0) ? t.roots[r] : t.nbr[h - 1][pslot];
            if (parent < 0) continue;
            for (int j = 0; j < f; ++j) {
                const int32_t c = cur[pslot * f + j];
                if (c < 0) continue;
                join_edge(t, 0, (uint32_t)c, (uint32_t)parent, out.edges);
                out.nodes.push_back((uint32_t)c);
            }
        }
        width_prev *= f;
    }
    out.nodes.push_back((uint32_t)t.roots[r]);
}

// array_distinct keeping first occurrences: one pass through a small open-addressing table (a neighbourhood has a few
// hundred entries; two std::sort calls per root were a third of the encoder's time at F = 128)
void distinct_nodes(RootPlan& out) {
    const size_t n = out.nodes.size();
    size_t cap = 16;
    while (cap < 2 * n) cap <<= 1;
    auto& tab = out.tmp;  // .first = key, .second = 1 when the slot is taken
    tab.assign(cap, {0u, 0u});
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t v = out.nodes[i];
        size_t h = ((size_t)v * 0x9E3779B1u) & (cap - 1);
        bool seen = false;
        while (tab[h].second) {
            if (tab[h].first == v) {
                seen = true;
                break;
            }
            h = (h + 1) & (cap - 1);
        }
        if (!seen) {
            tab[h] = {v, 1u};
            out.nodes[m++] = v;
        }
    }
    out.nodes.resize(m);
}

// array_distinct over Edge structs (src, dst, type, feature_values): two records of one (src, dst) pair collapse only
// when their feature rows are byte-identical.
void distinct_edges(const Tables& t, std::vector<EdgeRef>& e) {
    std::sort(e.begin(), e.end(), [](const EdgeRef& a, const EdgeRef& b) {
        return a.src != b.src ? a.src < b.src : a.dst != b.dst ? a.dst < b.dst : a.row < b.row;
    });
    size_t m = 0;
    for (size_t i = 0; i < e.size(); ++i) {
        bool dup = false;
        for (size_t k = m; k-- > 0 && e[k].src == e[i].src && e[k].dst == e[i].dst;) {
            const ETab& m = t.tab[0];
            if (e[k].row == e[i].row || m.Fe == 0 ||
                memcmp(m.ef + (size_t)e[k].row * m.Fe, m.ef + (size_t)e[i].row * m.Fe, sizeof(float) * (size_t)m.Fe) == 0) {
                dup = true;
                break;
            }
        }
        if (!dup) e[m++] = e[i];
    }
    e.resize(m);
}

void plan_root(const Tables& t, int64_t r, RootPlan& out) {
    out.nodes.clear();
    out.edges.clear();
    out.pos_edges.clear();
    out.neg_edges.clear();
    walk_tree(t, r, out);
    distinct_nodes(out);
}

// NodeAnchorBasedLinkPredictionSample of anchor roots[r]: neighbourhood = array_distinct(root's ++ every positive's ++
// every hard negative's) (lookupDstNodeNeighborhood + the merges at NodeAnchorBasedLinkPredictionTask.scala:186-209 and
// UserDefinedLabelsNodeAnchorBasedLinkPredictionTask.scala:405-581; a directed source-only anchor has an empty tree of
// its own and keeps the labels' neighbourhoods + itself, NodeAnchorBasedLinkPredictionBaseTask.scala:200-278),
// pos_edges / hard_neg_edges = hydrateTaskBasedEdges against the table the labels were sampled from (:280-334).
// Returns false if the anchor has no positive (no sample: the INNER JOINs drop it; negatives are LEFT JOINed).
bool plan_anchor(const Tables& t, int64_t r, int num_pos, const int32_t* pos, const int64_t* pos_tree, int num_neg, const int32_t* neg,
                 const int64_t* neg_tree, RootPlan& out) {
    out.nodes.clear();
    out.edges.clear();
    out.pos_edges.clear();
    out.neg_edges.clear();
    bool any = false;
    for (int j = 0; j < num_pos; ++j) any |= pos[r * num_pos + j] >= 0;
    if (!any) return false;
    walk_tree(t, r, out);
    const uint32_t root = (uint32_t)t.roots[r];
    for (int j = 0; j < num_pos; ++j) {
        const int32_t p = pos[r * num_pos + j];
        if (p < 0) continue;
        if (pos_tree[r * num_pos + j] >= 0) walk_tree(t, pos_tree[r * num_pos + j], out);
        else out.nodes.push_back((uint32_t)p);
        join_edge(t, 1, root, (uint32_t)p, out.pos_edges);  // (_src_node = root, _dst_node = positive)
    }
    if (out.pos_edges.empty()) return false;
    for (int j = 0; j < num_neg; ++j) {
        const int32_t q = neg[r * num_neg + j];
        if (q < 0) continue;
        if (neg_tree[r * num_neg + j] >= 0) walk_tree(t, neg_tree[r * num_neg + j], out);
        else out.nodes.push_back((uint32_t)q);
        join_edge(t, 2, root, (uint32_t)q, out.neg_edges);
    }
    // first-seen order with the root's own neighbourhood first
    distinct_nodes(out);
    distinct_edges(t, out.edges);
    return true;
}

struct Sizes {
    size_t root_node, graph, labels, pos, neg, message;
};

inline int ef_len(const Tables& t, const EdgeRef& e) { return (t.tab[e.tag].Fe > 0 && e.row >= 0) ? t.tab[e.tag].Fe : 0; }
inline const float* ef_row(const Tables& t, const EdgeRef& e) {
    return ef_len(t, e) ? t.tab[e.tag].ef + (size_t)e.row * t.tab[e.tag].Fe : nullptr;
}

Sizes message_sizes(const Tables& t, const RootPlan& pl, uint32_t root, bool with_label, int32_t label, size_t label_type_len) {
    Sizes s{};
    s.root_node = node_size(root, t.ntype, t.F);
    s.graph = 0;
    for (uint32_t v : pl.nodes) {
        const size_t ns = node_size(v, t.ntype, t.F);
        s.graph += 1 + varint_size(ns) + ns;
    }
    for (const auto& e : pl.edges) {
        const size_t es = edge_size(e.src, e.dst, t.etype, ef_len(t, e));
        s.graph += 1 + varint_size(es) + es;
    }
    s.labels = 0;
    if (with_label) {
        size_t ls = 0;
        if (label_type_len) ls += 1 + varint_size(label_type_len) + label_type_len;
        if (label != 0) ls += 1 + varint_size((uint64_t)(int64_t)label);  // int32: negative values sign-extend to 10 bytes
        s.labels = 1 + varint_size(ls) + ls;
    }
    s.pos = 0;
    for (const auto& e : pl.pos_edges) {
        const size_t es = edge_size(e.src, e.dst, t.etype, ef_len(t, e));
        s.pos += 1 + varint_size(es) + es;
    }
    s.neg = 0;
    for (const auto& e : pl.neg_edges) {
        const size_t es = edge_size(e.src, e.dst, t.etype, ef_len(t, e));
        s.neg += 1 + varint_size(es) + es;
    }
    s.message = 1 + varint_size(s.root_node) + s.root_node;
    if (s.graph > 0) s.message += 1 + varint_size(s.graph) + s.graph;
    s.message += s.labels + s.pos + s.neg;
    return s;
}

inline uint8_t* put_graph_body(const Tables& t, const RootPlan& pl, uint8_t* p) {
    // Graph { repeated Node nodes = 2; repeated Edge edges = 3; }
    for (uint32_t v : pl.nodes) {
        *p++ = 0x12;
        p = put_varint(p, node_size(v, t.ntype, t.F));
        p = put_node(p, v, t.ntype, t.F > 0 ? t.x + (size_t)v * t.F : nullptr, t.F);
    }
    for (const auto& e : pl.edges) {
        *p++ = 0x1A;
        p = put_varint(p, edge_size(e.src, e.dst, t.etype, ef_len(t, e)));
        p = put_edge(p, e.src, e.dst, t.etype, ef_row(t, e), ef_len(t, e));
    }
    return p;
}

int encode_samples(int32_t kind, int64_t n_roots, int64_t n_emit, const Tables& t, const int32_t* labels, const char* label_type,
                   int32_t num_pos, const int32_t* pos, const int64_t* pos_tree, int32_t num_neg, const int32_t* neg,
                   const int64_t* neg_tree, int32_t tfrecord_framing, uint8_t** out, int64_t* out_bytes, int64_t* record_offsets) {
    *out = nullptr;
    *out_bytes = 0;
    const size_t lt_len = label_type ? strlen(label_type) : 0;
    std::vector<int64_t> rec((size_t)n_emit + 1, 0);
    const int64_t no_label = INT32_MIN;
    // two passes over the roots: sizes, then bytes
    uint8_t* buf = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel
        {
            RootPlan pl;
#pragma omp for schedule(dynamic, 256)
            for (int64_t r = 0; r < n_emit; ++r) {
                if (pass == 1 && rec[(size_t)r + 1] == rec[(size_t)r]) continue;
                const uint32_t root = (uint32_t)t.roots[r];
                int32_t label = 0;
                if (kind == 1) {
                    label = labels[root];
                    if (label == no_label) continue;  // unlabeled node: no SupervisedNodeClassificationSample (inner join with the labels)
                }
                if (kind == 2) {
                    if (!plan_anchor(t, r, num_pos, pos, pos_tree, num_neg, neg, neg_tree, pl)) continue;
                } else {
                    plan_root(t, r, pl);
                }
                const Sizes s = message_sizes(t, pl, root, kind == 1, label, lt_len);
                if (pass == 0) {
                    rec[(size_t)r + 1] = (int64_t)s.message + (tfrecord_framing ? 16 : 0);
                    continue;
                }
                uint8_t* p = buf + rec[(size_t)r];
                uint8_t* payload = p;
                if (tfrecord_framing) {
                    const uint64_t len = s.message;
                    memcpy(p, &len, 8);
                    const uint32_t c = mask_crc(crc32c(p, 8));
                    memcpy(p + 8, &c, 4);
                    p += 12;
                    payload = p;
                }
                // root_node = 1
                *p++ = 0x0A;
                p = put_varint(p, s.root_node);
                p = put_node(p, root, t.ntype, t.F > 0 ? t.x + (size_t)root * t.F : nullptr, t.F);
                if (kind == 2) {
                    // hard_neg_edges = 2 (user-defined negatives only), neighborhood = 3, pos_edges = 4; neg_edges = 5 stays
                    // empty (castToTrainingSampleProtoSchema, NodeAnchorBasedLinkPredictionBaseTask.scala:388-426)
                    for (const auto& e : pl.neg_edges) {
                        *p++ = 0x12;
                        p = put_varint(p, edge_size(e.src, e.dst, t.etype, ef_len(t, e)));
                        p = put_edge(p, e.src, e.dst, t.etype, ef_row(t, e), ef_len(t, e));
                    }
                    if (s.graph > 0) {
                        *p++ = 0x1A;
                        p = put_varint(p, s.graph);
                        p = put_graph_body(t, pl, p);
                    }
                    for (const auto& e : pl.pos_edges) {
                        *p++ = 0x22;
                        p = put_varint(p, edge_size(e.src, e.dst, t.etype, ef_len(t, e)));
                        p = put_edge(p, e.src, e.dst, t.etype, ef_row(t, e), ef_len(t, e));
                    }
                } else {
                    // neighborhood = 2
                    if (s.graph > 0) {
                        *p++ = 0x12;
                        p = put_varint(p, s.graph);
                        p = put_graph_body(t, pl, p);
                    }
                }
                // root_node_labels = 3 : Label { string label_type = 1; int32 label = 2; }
                if (kind == 1) {
                    size_t ls = 0;
                    if (lt_len) ls += 1 + varint_size(lt_len) + lt_len;
                    if (label != 0) ls += 1 + varint_size((uint64_t)(int64_t)label);
                    *p++ = 0x1A;
                    p = put_varint(p, ls);
                    if (lt_len) {
                        *p++ = 0x0A;
                        p = put_varint(p, lt_len);
                        memcpy(p, label_type, lt_len);
                        p += lt_len;
                    }
                    if (label != 0) {
                        *p++ = 0x10;
                        p = put_varint(p, (uint64_t)(int64_t)label);
                    }
                }
                if (tfrecord_framing) {
                    const uint32_t c = mask_crc(crc32c(payload, (size_t)(p - payload)));
                    memcpy(p, &c, 4);
                    p += 4;
                }
            }
        }
        if (pass == 0) {
            for (int64_t r = 0; r < n_emit; ++r) rec[(size_t)r + 1] += rec[(size_t)r];
            const int64_t total = rec[(size_t)n_emit];
            buf = alloc_out((size_t)(total > 0 ? total : 0));
            if (!buf) return GIGL_E_NOMEM;
        }
    }
    if (record_offsets) memcpy(record_offsets, rec.data(), sizeof(int64_t) * ((size_t)n_emit + 1));
    *out = buf;
    *out_bytes = rec[(size_t)n_emit];
    return GIGL_OK;
}

}  // namespace

extern "C" {

uint32_t gigl_crc32c_masked(const void* data, int64_t n) { return mask_crc(crc32c((const uint8_t*)data, (size_t)n)); }

void gigl_free_host(void* p) { free(p); }

int gigl_encode_samples_host(int32_t kind, int64_t n_roots, const int32_t* roots, const int32_t* fanouts, int32_t n_hops,
                             const int32_t* const* nbr, const float* x, int32_t F, int32_t condensed_node_type,
                             int32_t condensed_edge_type, const int32_t* labels, const char* label_type, int32_t tfrecord_framing,
                             uint8_t** out, int64_t* out_bytes, int64_t* record_offsets) {
    if (kind != 0 && kind != 1) return GIGL_E_INVALID;
    return gigl_encode_samples_ex_host(kind, n_roots, n_roots, roots, fanouts, n_hops, nbr, x, F, condensed_node_type, condensed_edge_type,
                                       nullptr, nullptr, nullptr, nullptr, 0, labels, label_type, 0, nullptr, nullptr, tfrecord_framing,
                                       out, out_bytes, record_offsets);
}

static bool etab_ok(const gigl_edge_table* e) {
    if (!e) return true;
    if ((e->rowptr == nullptr) != (e->col == nullptr)) return false;
    return e->n_feat >= 0 && (e->n_feat == 0 || (e->feat && e->rowptr));
}
static ETab etab_of(const gigl_edge_table* e) {
    if (!e) return ETab{nullptr, nullptr, nullptr, nullptr, 0};
    return ETab{e->rowptr, e->col, e->edge_rows, e->feat, e->n_feat};
}
static bool labels_ok(int64_t n_emit, int64_t n_roots, const int32_t* roots, int32_t n, const int32_t* ids, const int64_t* trees) {
    for (int64_t i = 0; i < n_emit * n; ++i)
        if (trees[i] >= n_roots || (ids[i] >= 0 && trees[i] >= 0 && roots[trees[i]] != ids[i])) return false;
    return true;
}

int gigl_encode_samples_ex_host(int32_t kind, int64_t n_roots, int64_t n_emit, const int32_t* roots, const int32_t* fanouts,
                                int32_t n_hops, const int32_t* const* nbr, const float* x, int32_t F, int32_t condensed_node_type,
                                int32_t condensed_edge_type, const int64_t* rowptr, const int32_t* col, const int32_t* edge_rows,
                                const float* edge_feat, int32_t Fe, const int32_t* labels, const char* label_type, int32_t num_pos,
                                const int32_t* pos, const int64_t* pos_tree, int32_t tfrecord_framing, uint8_t** out,
                                int64_t* out_bytes, int64_t* record_offsets) {
    if (!out || !out_bytes || n_roots < 0 || n_emit < 0 || n_emit > n_roots || n_hops < 1 || n_hops > GIGL_MAX_HOPS || !fanouts || !nbr ||
        (n_roots > 0 && !roots))
        return GIGL_E_INVALID;
    if (kind < 0 || kind > 2) return GIGL_E_INVALID;
    if (kind == 1 && !labels) return GIGL_E_INVALID;
    if (kind == 2 && (num_pos < 1 || !pos || !pos_tree)) return GIGL_E_INVALID;
    if (F < 0 || (F > 0 && !x)) return GIGL_E_INVALID;
    const gigl_edge_table main_tab{rowptr, col, edge_rows, edge_feat, Fe};
    if (!etab_ok(&main_tab)) return GIGL_E_INVALID;
    if (kind == 2 && !labels_ok(n_emit, n_roots, roots, num_pos, pos, pos_tree)) return GIGL_E_INVALID;
    const ETab m = etab_of(&main_tab);
    const Tables t{roots, fanouts, n_hops, nbr, x, F, condensed_node_type, condensed_edge_type, {m, m, ETab{nullptr, nullptr, nullptr, nullptr, 0}}};
    return encode_samples(kind, n_roots, n_emit, t, labels, label_type, num_pos, pos, pos_tree, 0, nullptr, nullptr, tfrecord_framing, out,
                          out_bytes, record_offsets);
}

int gigl_encode_link_samples_host(int64_t n_roots, int64_t n_emit, const int32_t* roots, const int32_t* fanouts, int32_t n_hops,
                                  const int32_t* const* nbr, const float* x, int32_t F, int32_t condensed_node_type,
                                  int32_t condensed_edge_type, const gigl_edge_table* main_edges, const gigl_edge_table* pos_edges,
                                  const gigl_edge_table* neg_edges, int32_t num_pos, const int32_t* pos, const int64_t* pos_tree,
                                  int32_t num_neg, const int32_t* neg, const int64_t* neg_tree, int32_t tfrecord_framing, uint8_t** out,
                                  int64_t* out_bytes, int64_t* record_offsets) {
    if (!out || !out_bytes || n_roots < 0 || n_emit < 0 || n_emit > n_roots || n_hops < 1 || n_hops > GIGL_MAX_HOPS || !fanouts || !nbr ||
        (n_roots > 0 && !roots))
        return GIGL_E_INVALID;
    if (num_pos < 1 || !pos || !pos_tree || num_neg < 0 || (num_neg > 0 && (!neg || !neg_tree))) return GIGL_E_INVALID;
    if (F < 0 || (F > 0 && !x)) return GIGL_E_INVALID;
    if (!etab_ok(main_edges) || !etab_ok(pos_edges) || !etab_ok(neg_edges)) return GIGL_E_INVALID;
    if (!labels_ok(n_emit, n_roots, roots, num_pos, pos, pos_tree)) return GIGL_E_INVALID;
    if (num_neg > 0 && !labels_ok(n_emit, n_roots, roots, num_neg, neg, neg_tree)) return GIGL_E_INVALID;
    const ETab m = etab_of(main_edges);
    const Tables t{roots, fanouts, n_hops, nbr, x, F, condensed_node_type, condensed_edge_type,
                   {m, pos_edges ? etab_of(pos_edges) : m, etab_of(neg_edges)}};
    return encode_samples(2, n_roots, n_emit, t, nullptr, nullptr, num_pos, pos, pos_tree, num_neg, neg, neg_tree, tfrecord_framing, out,
                          out_bytes, record_offsets);
}

}  // extern "C"

// ---- typed (heterogeneous) samples from the ops of a SamplingOp DAG -------------------------------------------------
// GraphDBSampler.getKHopSubgraphForRootNode unions the ops' edge and node SETS and adds the root
// (scala_spark35/subgraph_sampler/src/main/scala/libs/sampler/GraphDBSampler.scala:129-148); nodes are then hydrated with
// their type's feature row and - when an edge type of the DAG carries features - edges LEFT JOIN the hydrated edge table
// on (_from, _to, _condensed_edge_type) (SGSTask.hydrateRnn, scala_spark35/.../libs/utils/SGSTask.scala:200-337).  The
// typed task's main sample merges the anchor's neighbourhood with its positives' by key
// (GraphDBNodeAnchorBasedLinkPredictionTask.scala:283-470, GraphPbWrappers.mergeGraphs :43-68).
namespace {

struct TNode {
    int32_t type;
    uint32_t id;
    bool operator<(const TNode& b) const { return type != b.type ? type < b.type : id < b.id; }
};
struct TEdge {
    int32_t type;
    uint32_t src, dst;
    int64_t row;  // feature row of the type's edge table, -1 = none
    bool operator<(const TEdge& b) const {
        return type != b.type ? type < b.type : src != b.src ? src < b.src : dst != b.dst ? dst < b.dst : row < b.row;
    }
};

struct TypedCtx {
    int32_t n_node_types;
    const gigl_node_table* node_tables;
    int32_t n_edge_types;
    const gigl_edge_table* edge_tables;  // nullptr: edges pass through un-hydrated

    const float* node_feat(int32_t type, uint32_t id, int& F) const {
        F = (type >= 0 && type < n_node_types) ? node_tables[type].n_feat : 0;
        return F > 0 ? node_tables[type].x + (size_t)id * F : nullptr;
    }
    const gigl_edge_table* etab(int32_t type) const {
        return (edge_tables && type >= 0 && type < n_edge_types && edge_tables[type].rowptr) ? &edge_tables[type] : nullptr;
    }
    int edge_feat_len(const TEdge& e) const {
        const gigl_edge_table* t = etab(e.type);
        return (t && t->n_feat > 0 && e.row >= 0) ? t->n_feat : 0;
    }
    const float* edge_feat(const TEdge& e) const {
        const int Fe = edge_feat_len(e);
        return Fe ? edge_tables[e.type].feat + (size_t)e.row * Fe : nullptr;
    }
    // LEFT JOIN of one (type, src, dst) key with the type's edge records: one Edge per matching record (all_records), or
    // the first one only (the by-key merge of mergeGraphs); a key without a record / without a table stays feature-less.
    void hydrate(bool enabled, const TEdge& key, bool all_records, std::vector<TEdge>& out) const {
        const gigl_edge_table* t = enabled ? etab(key.type) : nullptr;
        if (!t) {
            out.push_back({key.type, key.src, key.dst, -1});
            return;
        }
        const int32_t* b = t->col + t->rowptr[key.dst];
        const int32_t* en = t->col + t->rowptr[key.dst + 1];
        auto r = std::equal_range(b, en, (int32_t)key.src);
        if (r.first == r.second) {
            out.push_back({key.type, key.src, key.dst, -1});
            return;
        }
        for (const int32_t* q = r.first; q != r.second; ++q) {
            const int64_t slot = q - t->col;
            out.push_back({key.type, key.src, key.dst, t->n_feat > 0 ? (t->edge_rows ? (int64_t)t->edge_rows[slot] : slot) : -1});
            if (!all_records) break;
        }
    }
};

bool dag_tree_ok(const gigl_dag_tree* t, std::vector<int64_t>& width) {
    if (!t || t->n_roots < 0 || (t->n_roots > 0 && !t->roots) || t->n_ops < 0 || (t->n_ops > 0 && !t->ops) || t->root_node_type < 0) return false;
    width.assign((size_t)t->n_ops, 1);
    for (int o = 0; o < t->n_ops; ++o) {
        const gigl_dag_op& op = t->ops[o];
        if (op.parent >= o || op.parent < -1 || op.fanout < 1 || op.fanout > GIGL_MAX_FANOUT || (t->n_roots > 0 && !op.nbr) ||
            op.result_node_type < 0 || op.condensed_edge_type < -1)
            return false;  // ops come in topological order
        width[(size_t)o] = (op.parent < 0 ? 1 : width[(size_t)op.parent]) * op.fanout;
    }
    return true;
}

// appends the (non-distinct) typed nodes and edge keys of root r's sampled DAG, the root included
void walk_dag(const gigl_dag_tree& t, const std::vector<int64_t>& width, int64_t r, std::vector<TNode>& nodes, std::vector<TEdge>& edges) {
    nodes.push_back({t.root_node_type, (uint32_t)t.roots[r]});
    for (int o = 0; o < t.n_ops; ++o) {
        const gigl_dag_op& op = t.ops[o];
        const int64_t w = width[(size_t)o];
        for (int64_t s_ = r * w; s_ < (r + 1) * w; ++s_) {
            const int32_t c = op.nbr[s_];
            if (c < 0) continue;
            const int32_t par = op.parent < 0 ? t.roots[r] : t.ops[op.parent].nbr[s_ / op.fanout];
            if (par < 0) continue;
            nodes.push_back({op.result_node_type, (uint32_t)c});
            edges.push_back(op.outgoing ? TEdge{op.condensed_edge_type, (uint32_t)par, (uint32_t)c, -1}
                                        : TEdge{op.condensed_edge_type, (uint32_t)c, (uint32_t)par, -1});
        }
    }
}

template <class T>
void sort_unique(std::vector<T>& v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end(), [](const T& a, const T& b) { return !(a < b) && !(b < a); }), v.end());
}

int encode_typed(int32_t kind, const gigl_dag_tree* anchors, const gigl_dag_tree* targets, int32_t num_pos, const int32_t* pos,
                 const int64_t* pos_tree, int32_t pos_edge_type, int32_t include_isolated, int32_t hydrate_flags, const TypedCtx& tc,
                 int32_t tfrecord_framing, uint8_t** out, int64_t* out_bytes, int64_t* record_offsets) {
    std::vector<int64_t> aw, tw;
    if (!out || !out_bytes || !dag_tree_ok(anchors, aw) || (kind != 0 && kind != 2)) return GIGL_E_INVALID;
    if (kind == 2 && (num_pos < 0 || (num_pos > 0 && anchors->n_roots > 0 && (!pos || !pos_tree)) || !dag_tree_ok(targets, tw)))
        return GIGL_E_INVALID;
    const int64_t n_roots = anchors->n_roots;
    if (kind == 2)
        for (int64_t i = 0; i < n_roots * num_pos; ++i)
            if (pos[i] >= 0 && (pos_tree[i] < -1 || pos_tree[i] >= targets->n_roots || (pos_tree[i] >= 0 && targets->roots[pos_tree[i]] != pos[i])))
                return GIGL_E_INVALID;
    *out = nullptr;
    *out_bytes = 0;
    const bool hyd_graph = hydrate_flags & 1, hyd_pos = hydrate_flags & 2;
    std::vector<int64_t> rec((size_t)n_roots + 1, 0);
    uint8_t* buf = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
#pragma omp parallel
        {
            std::vector<TNode> nodes;
            std::vector<TEdge> keys, edges, pos_keys, pos_edges;
#pragma omp for schedule(dynamic, 256)
            for (int64_t r = 0; r < n_roots; ++r) {
                if (pass == 1 && rec[(size_t)r + 1] == rec[(size_t)r]) continue;
                nodes.clear();
                keys.clear();
                edges.clear();
                pos_keys.clear();
                pos_edges.clear();
                const uint32_t root = (uint32_t)anchors->roots[r];
                const int32_t rtype = anchors->root_node_type;
                if (kind == 2) {
                    for (int j = 0; j < num_pos; ++j)
                        if (pos[r * num_pos + j] >= 0) pos_keys.push_back({pos_edge_type, root, (uint32_t)pos[r * num_pos + j], -1});
                    if (pos_keys.empty() && !include_isolated) continue;  // INNER JOIN with the positives (:404-420)
                }
                walk_dag(*anchors, aw, r, nodes, keys);
                if (kind == 2) {
                    for (int j = 0; j < num_pos; ++j) {
                        const int32_t pnode = pos[r * num_pos + j];
                        if (pnode < 0) continue;
                        const int64_t tr = pos_tree[r * num_pos + j];
                        if (tr >= 0) walk_dag(*targets, tw, tr, nodes, keys);
                        else nodes.push_back({targets->root_node_type, (uint32_t)pnode});
                    }
                    sort_unique(pos_keys);  // the query result is parsed into a SET of edges (GraphDBSampler.scala:196-205)
                    for (const TEdge& k : pos_keys) tc.hydrate(hyd_pos, k, true, pos_edges);
                }
                sort_unique(nodes);
                sort_unique(keys);
                // RootedNodeNeighborhood: collect_list over the LEFT JOIN = one Edge per record; main sample: merged by key
                for (const TEdge& k : keys) tc.hydrate(hyd_graph, k, kind == 0, edges);
                int Fr = 0;
                const float* xr = tc.node_feat(rtype, root, Fr);
                const size_t root_sz = node_size(root, rtype, Fr);
                size_t graph = 0, pos_sz = 0;
                for (const TNode& v : nodes) {
                    int F = 0;
                    tc.node_feat(v.type, v.id, F);
                    const size_t ns = node_size(v.id, v.type, F);
                    graph += 1 + varint_size(ns) + ns;
                }
                for (const TEdge& e : edges) {
                    const size_t es = edge_size(e.src, e.dst, e.type, tc.edge_feat_len(e));
                    graph += 1 + varint_size(es) + es;
                }
                for (const TEdge& e : pos_edges) {
                    const size_t es = edge_size(e.src, e.dst, e.type, tc.edge_feat_len(e));
                    pos_sz += 1 + varint_size(es) + es;
                }
                const size_t message = 1 + varint_size(root_sz) + root_sz + 1 + varint_size(graph) + graph + pos_sz;
                if (pass == 0) {
                    rec[(size_t)r + 1] = (int64_t)message + (tfrecord_framing ? 16 : 0);
                    continue;
                }
                uint8_t* p = buf + rec[(size_t)r];
                uint8_t* payload = p;
                if (tfrecord_framing) {
                    const uint64_t len = message;
                    memcpy(p, &len, 8);
                    const uint32_t c = mask_crc(crc32c(p, 8));
                    memcpy(p + 8, &c, 4);
                    p += 12;
                    payload = p;
                }
                *p++ = 0x0A;  // root_node = 1
                p = put_varint(p, root_sz);
                p = put_node(p, root, rtype, xr, Fr);
                // neighborhood: field 2 of RootedNodeNeighborhood, field 3 of NodeAnchorBasedLinkPredictionSample
                *p++ = kind == 2 ? 0x1A : 0x12;
                p = put_varint(p, graph);
                for (const TNode& v : nodes) {  // Graph { nodes = 2, edges = 3 }
                    int F = 0;
                    const float* xv = tc.node_feat(v.type, v.id, F);
                    *p++ = 0x12;
                    p = put_varint(p, node_size(v.id, v.type, F));
                    p = put_node(p, v.id, v.type, xv, F);
                }
                for (const TEdge& e : edges) {
                    const int Fe = tc.edge_feat_len(e);
                    *p++ = 0x1A;
                    p = put_varint(p, edge_size(e.src, e.dst, e.type, Fe));
                    p = put_edge(p, e.src, e.dst, e.type, tc.edge_feat(e), Fe);
                }
                for (const TEdge& e : pos_edges) {  // pos_edges = 4
                    const int Fe = tc.edge_feat_len(e);
                    *p++ = 0x22;
                    p = put_varint(p, edge_size(e.src, e.dst, e.type, Fe));
                    p = put_edge(p, e.src, e.dst, e.type, tc.edge_feat(e), Fe);
                }
                if (tfrecord_framing) {
                    const uint32_t c = mask_crc(crc32c(payload, (size_t)(p - payload)));
                    memcpy(p, &c, 4);
                    p += 4;
                }
            }
        }
        if (pass == 0) {
            for (int64_t r = 0; r < n_roots; ++r) rec[(size_t)r + 1] += rec[(size_t)r];
            buf = alloc_out((size_t)(rec[(size_t)n_roots] > 0 ? rec[(size_t)n_roots] : 0));
            if (!buf) return GIGL_E_NOMEM;
        }
    }
    if (record_offsets) memcpy(record_offsets, rec.data(), sizeof(int64_t) * ((size_t)n_roots + 1));
    *out = buf;
    *out_bytes = rec[(size_t)n_roots];
    return GIGL_OK;
}

bool typed_tables_ok(int32_t n_node_types, const gigl_node_table* node_tables, int32_t n_edge_types, const gigl_edge_table* edge_tables) {
    if (n_node_types < 0 || (n_node_types > 0 && !node_tables) || n_edge_types < 0) return false;
    for (int t = 0; t < n_node_types; ++t)
        if (node_tables[t].n_feat < 0 || (node_tables[t].n_feat > 0 && !node_tables[t].x)) return false;
    if (edge_tables)
        for (int t = 0; t < n_edge_types; ++t)
            if (edge_tables[t].rowptr && (!edge_tables[t].col || edge_tables[t].n_feat < 0 || (edge_tables[t].n_feat > 0 && !edge_tables[t].feat)))
                return false;
    return true;
}

}  // namespace

extern "C" {

int gigl_encode_dag_samples_host(int64_t n_roots, const int32_t* roots, int32_t root_node_type, int32_t n_ops, const gigl_dag_op* ops,
                                 int32_t n_node_types, const gigl_node_table* node_tables, int32_t tfrecord_framing, uint8_t** out,
                                 int64_t* out_bytes, int64_t* record_offsets) {
    if (!typed_tables_ok(n_node_types, node_tables, 0, nullptr)) return GIGL_E_INVALID;
    const gigl_dag_tree tree{n_roots, roots, root_node_type, n_ops, ops};
    const TypedCtx tc{n_node_types, node_tables, 0, nullptr};
    return encode_typed(0, &tree, nullptr, 0, nullptr, nullptr, -1, 0, 0, tc, tfrecord_framing, out, out_bytes, record_offsets);
}

int gigl_encode_typed_samples_host(int32_t kind, const gigl_dag_tree* anchors, const gigl_dag_tree* targets, int32_t num_pos,
                                   const int32_t* pos, const int64_t* pos_tree, int32_t pos_condensed_edge_type, int32_t include_isolated,
                                   int32_t hydrate_flags, int32_t n_node_types, const gigl_node_table* node_tables, int32_t n_edge_types,
                                   const gigl_edge_table* edge_tables, int32_t tfrecord_framing, uint8_t** out, int64_t* out_bytes,
                                   int64_t* record_offsets) {
    if (!typed_tables_ok(n_node_types, node_tables, n_edge_types, edge_tables)) return GIGL_E_INVALID;
    const TypedCtx tc{n_node_types, node_tables, n_edge_types, edge_tables};
    return encode_typed(kind, anchors, targets, num_pos, pos, pos_tree, pos_condensed_edge_type, include_isolated, hydrate_flags, tc,
                        tfrecord_framing, out, out_bytes, record_offsets);
}

// ---- TFRecord reading + tf.Example decoding ----------------------------------------------------------
// Splits a TFRecord byte stream into records; verifies both checksums when verify != 0.
// offsets / lengths: caller arrays of capacity max_records; returns the number of records, or < 0.
int64_t gigl_tfrecord_index_host(const uint8_t* data, int64_t n_bytes, int32_t verify, int64_t* offsets, int64_t* lengths,
                                 int64_t max_records) {
    if (!data || n_bytes < 0) return GIGL_E_INVALID;
    int64_t pos = 0, n = 0;
    while (pos < n_bytes) {
        if (n_bytes - pos < 16) return GIGL_E_INVALID;  // length (8) + its crc (4) + the payload crc (4): a truncated tail
        uint64_t len;
        uint32_t c;
        memcpy(&len, data + pos, 8);
        memcpy(&c, data + pos + 8, 4);
        if (verify && c != mask_crc(crc32c(data + pos, 8))) return GIGL_E_INVALID;
        if (len > (uint64_t)(n_bytes - pos - 16)) return GIGL_E_INVALID;
        if (verify) {
            memcpy(&c, data + pos + 12 + len, 4);
            if (c != mask_crc(crc32c(data + pos + 12, (size_t)len))) return GIGL_E_INVALID;
        }
        if (offsets && lengths) {
            if (n >= max_records) return GIGL_E_OVERFLOW;
            offsets[n] = pos + 12;
            lengths[n] = (int64_t)len;
        }
        ++n;
        pos += 16 + (int64_t)len;
    }
    return n;
}

// ---- TaskOutputValidator (scala/subgraph_sampler/src/main/scala/libs/task/TaskOutputValidator.scala:29-108) ------------
// Every sample the component is about to write is parsed back from its OWN bytes and held to the reference's check: both
// endpoints of every neighbourhood edge - and, for NodeAnchorBasedLinkPredictionSample, of every pos / neg / hard-neg edge -
// are among the neighbourhood's nodes, compared as (node id, condensed node type) with the types an edge's condensed edge
// type implies (GraphMetadataPbWrapper.getFeaturelessNodePbsFromEdge :75-83); a sample without a neighbourhood fails.
namespace {
struct PbField {
    uint32_t num;
    uint32_t wire;
    uint64_t val;          // varint / fixed value
    const uint8_t* data;   // length-delimited payload
    uint64_t len;
};
// next field of a message; false at the end or on malformed input (ok = false)
inline bool pb_next(const uint8_t*& p, const uint8_t* end, PbField& f, bool& ok) {
    if (p >= end) return false;
    uint64_t key;
    if (!get_varint(p, end, key)) {
        ok = false;
        return false;
    }
    f.num = (uint32_t)(key >> 3);
    f.wire = (uint32_t)(key & 7);
    f.val = 0;
    f.data = nullptr;
    f.len = 0;
    switch (f.wire) {
        case 0:
            if (!get_varint(p, end, f.val)) ok = false;
            break;
        case 1:
            if (end - p < 8) ok = false; else p += 8;
            break;
        case 5:
            if (end - p < 4) ok = false; else p += 4;
            break;
        case 2:
            if (!get_varint(p, end, f.len) || f.len > (uint64_t)(end - p)) {
                ok = false;
            } else {
                f.data = p;
                p += f.len;
            }
            break;
        default:
            ok = false;
    }
    return ok;
}
inline uint64_t node_key(uint64_t id, uint64_t type) { return (type << 32) | (id & 0xffffffffULL); }
struct EdgeEnds {
    uint64_t src = 0, dst = 0, type = 0;
};
inline bool parse_edge(const uint8_t* p, const uint8_t* end, EdgeEnds& e) {
    bool ok = true;
    PbField f;
    while (pb_next(p, end, f, ok)) {
        if (f.num == 1 && f.wire == 0) e.src = f.val;
        if (f.num == 2 && f.wire == 0) e.dst = f.val;
        if (f.num == 3 && f.wire == 0) e.type = f.val;
    }
    return ok;
}
// 0 = valid, 1 = malformed bytes, 2 = no neighbourhood, 3 = an edge endpoint is not among the neighbourhood nodes
int validate_sample(const uint8_t* p, const uint8_t* end, int32_t kind, int32_t n_edge_types, const int32_t* src_type,
                    const int32_t* dst_type, std::vector<uint64_t>& nodes, std::vector<EdgeEnds>& edges) {
    nodes.clear();
    edges.clear();
    bool ok = true, has_graph = false;
    PbField f;
    const uint32_t graph_field = (kind == 2) ? 3u : 2u;  // training_samples_schema.proto: neighborhood = 2 (RNN / SNC), 3 (NABLP)
    while (pb_next(p, end, f, ok)) {
        if (f.wire != 2) continue;
        if (f.num == graph_field) {
            has_graph = true;
            const uint8_t* q = f.data;
            const uint8_t* qe = f.data + f.len;
            PbField g;
            while (pb_next(q, qe, g, ok)) {
                if (g.wire != 2) continue;
                if (g.num == 2) {  // Graph.nodes
                    const uint8_t* r = g.data;
                    const uint8_t* re = g.data + g.len;
                    uint64_t id = 0, type = 0;
                    PbField h;
                    while (pb_next(r, re, h, ok)) {
                        if (h.num == 1 && h.wire == 0) id = h.val;
                        if (h.num == 2 && h.wire == 0) type = h.val;
                    }
                    nodes.push_back(node_key(id, type));
                } else if (g.num == 3) {  // Graph.edges
                    EdgeEnds e;
                    if (!parse_edge(g.data, g.data + g.len, e)) ok = false;
                    edges.push_back(e);
                }
            }
        } else if (kind == 2 && (f.num == 2 || f.num == 4 || f.num == 5)) {  // hard_neg_edges / pos_edges / neg_edges
            EdgeEnds e;
            if (!parse_edge(f.data, f.data + f.len, e)) ok = false;
            edges.push_back(e);
        }
    }
    if (!ok) return 1;
    if (!has_graph) return 2;
    std::sort(nodes.begin(), nodes.end());
    for (const EdgeEnds& e : edges) {
        uint64_t st = 0, dt = 0;  // DefaultCondensedNodeType
        if (src_type && dst_type) {
            if (e.type >= (uint64_t)n_edge_types) return 3;
            st = (uint64_t)(uint32_t)src_type[e.type];
            dt = (uint64_t)(uint32_t)dst_type[e.type];
        }
        if (!std::binary_search(nodes.begin(), nodes.end(), node_key(e.src, st)) ||
            !std::binary_search(nodes.begin(), nodes.end(), node_key(e.dst, dt)))
            return 3;
    }
    return 0;
}
}  // namespace

// data: the encoder's output (TFRecord-framed when tfrecord_framing != 0, else n_records payloads given by offsets /
// lengths).  kind: 0 = RootedNodeNeighborhood, 1 = SupervisedNodeClassificationSample, 2 = NodeAnchorBasedLinkPredictionSample.
// edge_src_type / edge_dst_type: condensed node types of every condensed edge type's endpoints (NULL = homogeneous: every
// node type 0).  Returns GIGL_OK, or GIGL_E_INVALID with *bad_record = the first offending record and *reason = 1 malformed,
// 2 neighbourhood missing, 3 edge endpoint outside the neighbourhood nodes.
int gigl_validate_samples_host(const uint8_t* data, int64_t n_bytes, int32_t kind, int32_t n_edge_types, const int32_t* edge_src_type,
                               const int32_t* edge_dst_type, int64_t* n_records_out, int64_t* bad_record, int32_t* reason) {
    if (!data || n_bytes < 0 || kind < 0 || kind > 2 || ((edge_src_type == nullptr) != (edge_dst_type == nullptr))) return GIGL_E_INVALID;
    const int64_t n = gigl_tfrecord_index_host(data, n_bytes, 0, nullptr, nullptr, 0);
    if (n < 0) return (int)n;
    std::vector<int64_t> off((size_t)(n > 0 ? n : 1)), len((size_t)(n > 0 ? n : 1));
    if (n > 0 && gigl_tfrecord_index_host(data, n_bytes, 0, off.data(), len.data(), n) != n) return GIGL_E_INVALID;
    int64_t first_bad = n;
    int32_t why = 0;
#pragma omp parallel
    {
        std::vector<uint64_t> nodes;
        std::vector<EdgeEnds> edges;
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            const int r = validate_sample(data + off[i], data + off[i] + len[i], kind, n_edge_types, edge_src_type, edge_dst_type, nodes, edges);
            if (r != 0) {
#pragma omp critical
                if (i < first_bad) {
                    first_bad = i;
                    why = r;
                }
            }
        }
    }
    if (n_records_out) *n_records_out = n;
    if (first_bad < n) {
        if (bad_record) *bad_record = first_bad;
        if (reason) *reason = why;
        return GIGL_E_INVALID;
    }
    return GIGL_OK;
}

// Decodes one named feature of every tf.Example record into a dense column.
// dtype: 0 = int64 (Int64List, written to out_i64), 1 = float (FloatList -> out_f32; an Int64List is cast, which is what
// `cast(col as array<float>)` does in loadNodeDataframeIntoSparkSql :90-104).  width = values per record (records with a
// different count are an error).  Returns GIGL_OK, or GIGL_E_RANGE if a record lacks the feature.
int gigl_examples_column_host(const uint8_t* data, int64_t n_records, const int64_t* offsets, const int64_t* lengths,
                              const char* name, int32_t dtype, int32_t width, int64_t* out_i64, float* out_f32) {
    if (!data || !offsets || !lengths || !name || width < 1 || (dtype == 0 && !out_i64) || (dtype == 1 && !out_f32) || dtype < 0 || dtype > 1)
        return GIGL_E_INVALID;
    const size_t name_len = strlen(name);
    int rc_all = GIGL_OK;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_records; ++r) {
        const uint8_t* p = data + offsets[r];
        const uint8_t* end = p + lengths[r];
        bool found = false;
        int rc = GIGL_OK;
        // Example { Features features = 1; }  Features { map<string, Feature> feature = 1; }
        while (p < end && rc == GIGL_OK && !found) {
            uint64_t tag, len;
            if (!get_varint(p, end, tag)) { rc = GIGL_E_INVALID; break; }
            if ((tag & 7) != 2) { rc = GIGL_E_INVALID; break; }
            if (!get_varint(p, end, len) || len > (uint64_t)(end - p)) { rc = GIGL_E_INVALID; break; }
            const uint8_t* fend = p + len;
            if ((tag >> 3) != 1) { p = fend; continue; }
            const uint8_t* q = p;  // Features
            while (q < fend && !found) {
                uint64_t t2, l2;
                if (!get_varint(q, fend, t2) || (t2 & 7) != 2 || !get_varint(q, fend, l2) || l2 > (uint64_t)(fend - q)) { rc = GIGL_E_INVALID; break; }
                const uint8_t* eend = q + l2;  // one map entry { string key = 1; Feature value = 2; }
                const uint8_t* e = q;
                const uint8_t* key = nullptr;
                size_t key_len = 0;
                const uint8_t* val = nullptr;
                size_t val_len = 0;
                while (e < eend) {
                    uint64_t t3, l3;
                    if (!get_varint(e, eend, t3) || (t3 & 7) != 2 || !get_varint(e, eend, l3) || l3 > (uint64_t)(eend - e)) { rc = GIGL_E_INVALID; break; }
                    if ((t3 >> 3) == 1) { key = e; key_len = (size_t)l3; }
                    if ((t3 >> 3) == 2) { val = e; val_len = (size_t)l3; }
                    e += l3;
                }
                if (rc != GIGL_OK) break;
                if (key && key_len == name_len && memcmp(key, name, name_len) == 0) {
                    found = true;
                    // Feature { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
                    const uint8_t* v = val;
                    const uint8_t* vend = val + val_len;
                    int count = 0;
                    while (v && v < vend && rc == GIGL_OK) {
                        uint64_t t4, l4;
                        if (!get_varint(v, vend, t4) || (t4 & 7) != 2 || !get_varint(v, vend, l4) || l4 > (uint64_t)(vend - v)) { rc = GIGL_E_INVALID; break; }
                        const uint8_t* lend = v + l4;
                        const int kindf = (int)(t4 >> 3);
                        // FloatList { repeated float value = 1 [packed] }  Int64List { repeated int64 value = 1 [packed] }
                        const uint8_t* w = v;
                        while (w < lend && rc == GIGL_OK) {
                            uint64_t t5;
                            if (!get_varint(w, lend, t5)) { rc = GIGL_E_INVALID; break; }
                            if (kindf == 2) {
                                if ((t5 & 7) == 2) {  // packed floats
                                    uint64_t l5;
                                    if (!get_varint(w, lend, l5) || l5 > (uint64_t)(lend - w) || (l5 & 3)) { rc = GIGL_E_INVALID; break; }
                                    for (uint64_t i = 0; i < l5; i += 4) {
                                        float fv;
                                        memcpy(&fv, w + i, 4);
                                        if (count < width) { if (dtype == 1) out_f32[r * width + count] = fv; else out_i64[r * width + count] = (int64_t)fv; }
                                        ++count;
                                    }
                                    w += l5;
                                } else if ((t5 & 7) == 5) {
                                    if (lend - w < 4) { rc = GIGL_E_INVALID; break; }
                                    float fv;
                                    memcpy(&fv, w, 4);
                                    w += 4;
                                    if (count < width) { if (dtype == 1) out_f32[r * width + count] = fv; else out_i64[r * width + count] = (int64_t)fv; }
                                    ++count;
                                } else { rc = GIGL_E_INVALID; }
                            } else if (kindf == 3) {
                                if ((t5 & 7) == 2) {  // packed varints
                                    uint64_t l5;
                                    if (!get_varint(w, lend, l5) || l5 > (uint64_t)(lend - w)) { rc = GIGL_E_INVALID; break; }
                                    const uint8_t* pend = w + l5;
                                    while (w < pend) {
                                        uint64_t iv;
                                        if (!get_varint(w, pend, iv)) { rc = GIGL_E_INVALID; break; }
                                        if (count < width) { if (dtype == 0) out_i64[r * width + count] = (int64_t)iv; else out_f32[r * width + count] = (float)(int64_t)iv; }
                                        ++count;
                                    }
                                } else if ((t5 & 7) == 0) {
                                    uint64_t iv;
                                    if (!get_varint(w, lend, iv)) { rc = GIGL_E_INVALID; break; }
                                    if (count < width) { if (dtype == 0) out_i64[r * width + count] = (int64_t)iv; else out_f32[r * width + count] = (float)(int64_t)iv; }
                                    ++count;
                                } else { rc = GIGL_E_INVALID; }
                            } else {
                                rc = GIGL_E_INVALID;  // bytes_list columns are not numeric
                            }
                        }
                        v = lend;
                    }
                    if (rc == GIGL_OK && count != width) rc = GIGL_E_INVALID;
                }
                q = eend;
            }
            p = fend;
        }
        if (rc == GIGL_OK && !found) rc = GIGL_E_RANGE;
        if (rc != GIGL_OK) {
#pragma omp critical
            rc_all = rc;
        }
    }
    return rc_all;
}

}  // extern "C"
