"""Small deterministic inputs for every sample type the host encoder writes.  Shared by tests/golden/make_golden.py
(which pushes the encoder's bytes through the REFERENCE's generated protobuf classes and commits what they see) and by
tests/test_encoder_vs_reference_pb2.py (which holds the encoder and the repo's parser to that view)."""
import numpy as np

from gigl_b200 import sample_io as sio
from oracle import oracle as O


def _edge_rows(src, dst, n):
    """directed graphs: CSR slot -> input record, slots ordered by (dst, src, record)"""
    idx = np.arange(len(src))
    return idx[np.lexsort((idx, src, dst))].astype(np.int32)


def cases():
    """{name: (kind, bytes)} with kind in rnn / snc / nablp."""
    rng = np.random.default_rng(77)
    n, e = 24, 90
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    src[80:], dst[80:] = src[:10], dst[:10]  # duplicate edge records with their own feature rows
    ef = rng.standard_normal((e, 2)).astype(np.float32)
    x = rng.standard_normal((n, 3)).astype(np.float32)
    rowptr, col = O.np_build_in_csr(src, dst, n, True)
    orow, ocol = O.np_build_in_csr(dst, src, n, True)
    rows = _edge_rows(src, dst, n)
    roots = np.arange(n, dtype=np.int32)
    fan = [3, 2]
    nbr, _ = O.c_sample_khop(rowptr, col, roots, fan)
    out = {}
    out["rnn_with_edge_features"] = ("rnn", sio.encode_samples(roots, fan, nbr, x, kind="rnn", condensed_node_type=0, condensed_edge_type=0,
                                                                csr=(rowptr, col), edge_rows=rows, edge_feat=ef)[0])
    out["rnn_no_features_types_unset"] = ("rnn", sio.encode_samples(roots, fan, nbr, None, kind="rnn", condensed_node_type=-1,
                                                                     condensed_edge_type=-1)[0])
    labels = np.arange(n, dtype=np.int32) - 5  # negative, zero and positive labels (int32 on the wire: negatives take 10 bytes)
    labels[3] = sio.INT32_MIN                  # unlabeled node: no sample
    out["snc_labels"] = ("snc", sio.encode_samples(roots, fan, nbr, x, kind="snc", labels=labels, label_type="node_label")[0])
    positives = O.np_sample_positives(orow, ocol, roots, 2)
    pos = np.full((n, 2), -1, np.int32)
    for u, ps in positives.items():
        pos[u, :len(ps)] = ps
    tree = np.where(pos >= 0, pos, -1).astype(np.int64)
    out["nablp_main_edges"] = ("nablp", sio.encode_samples(roots, fan, nbr, x, kind="nablp", csr=(rowptr, col), edge_rows=rows, edge_feat=ef,
                                                           pos=pos, pos_tree=tree)[0])
    # user-defined labels: positives / hard negatives from their own tables
    ps_, pd_ = rng.integers(0, n, 30), rng.integers(0, n, 30)
    ns_, nd_ = rng.integers(0, n, 30), rng.integers(0, n, 30)
    pf, nf = rng.standard_normal((30, 1)).astype(np.float32), rng.standard_normal((30, 1)).astype(np.float32)
    p_out, n_out = O.np_build_in_csr(pd_, ps_, n, True), O.np_build_in_csr(nd_, ns_, n, True)
    upos, uneg = O.np_sample_positives(p_out[0], p_out[1], roots, 2, call_no=3), O.np_sample_positives(n_out[0], n_out[1], roots, 2, call_no=4)

    def dense(d):
        a = np.full((n, 2), -1, np.int32)
        for u, lst in d.items():
            a[u, :len(lst)] = lst
        return a

    dp, dn = dense(upos), dense(uneg)
    main_tab = sio.HostEdgeTable((rowptr, col), rows, ef)
    pos_tab = sio.HostEdgeTable(O.np_build_in_csr(ps_, pd_, n, True), _edge_rows(ps_, pd_, n), pf)
    neg_tab = sio.HostEdgeTable(O.np_build_in_csr(ns_, nd_, n, True), _edge_rows(ns_, nd_, n), nf)
    out["nablp_user_defined_labels"] = ("nablp", sio.encode_link_samples(roots, fan, nbr, x, n, dp, np.where(dp >= 0, dp, -1).astype(np.int64),
                                                                          main_tab, pos_tab, dn, np.where(dn >= 0, dn, -1).astype(np.int64),
                                                                          neg_tab)[0])
    # typed: user / item graph, two ops, edge features on one type, positives over the featured type
    n_u, n_i = 12, 9
    nn = max(n_u, n_i)
    f_s, f_d = rng.integers(0, n_u, 40), rng.integers(0, n_u, 40)            # type 0: user -> user
    c_s, c_d = rng.integers(0, 8, 40), rng.integers(0, n_i, 40)              # type 1: user -> item, featured; users 8..11 never click
    cf = rng.standard_normal((40, 2)).astype(np.float32)
    xu, xi = rng.standard_normal((nn, 2)).astype(np.float32), rng.standard_normal((nn, 4)).astype(np.float32)
    inc0, inc1 = O.np_build_in_csr(f_s, f_d, nn, True), O.np_build_in_csr(c_s, c_d, nn, True)
    out1 = O.np_build_in_csr(c_d, c_s, nn, True)
    users = np.arange(n_u, dtype=np.int32)
    h1, _ = O.np_sample_chain([inc0], users, [2], [1])
    h2, _ = O.np_sample_chain([inc0, out1], users, [2, 2], [1, 2])
    uops = [dict(parent=-1, fanout=2, condensed_edge_type=0, result_node_type=0, nbr=h1[0]),
            dict(parent=0, fanout=2, condensed_edge_type=1, result_node_type=1, outgoing=True, nbr=h2[1])]
    tabs = [sio.HostEdgeTable(inc0, None, None), sio.HostEdgeTable(inc1, _edge_rows(c_s, c_d, nn), cf)]
    out["typed_rnn"] = ("rnn", sio.encode_typed_samples(users, 0, uops, [xu, xi], tabs, kind="rnn")[0])
    tpos, _ = O.np_sample_chain([out1], users, [1], [3])
    tpos = tpos[0].reshape(n_u, 1)
    items = np.unique(tpos[tpos >= 0]).astype(np.int32)
    t1, _ = O.np_sample_chain([inc1], items, [2], [1])
    tops = [dict(parent=-1, fanout=2, condensed_edge_type=1, result_node_type=0, nbr=t1[0])]
    ttree = np.where(tpos >= 0, np.searchsorted(items, np.maximum(tpos, 0)), -1).astype(np.int64)
    out["typed_nablp"] = ("nablp", sio.encode_typed_samples(users, 0, uops, [xu, xi], tabs, kind="nablp", pos=tpos, pos_tree=ttree,
                                                            pos_condensed_edge_type=1, target_roots=items, target_node_type=1, target_ops=tops,
                                                            include_isolated=True)[0])
    return out
