"""CPU: the sampler component's HOST logic (config parsing, batching, label / positive handling, hydration, zero-copy
TFRecord writing, sharding over ranks) without a GPU.  The CUDA sampler is replaced by an oracle-backed stand-in with the
same methods, and the bodies of the GPU component tests (tests/test_gpu_cli.py) are run unchanged; on the GPU box those
same bodies run against the real kernels.  Nothing here says anything about the kernels - the sampled index sets are the
oracle's by construction - it keeps the Python / C++ host side honest on every CPU run."""
import numpy as np
import pytest

import test_gpu_cli as G
from gigl_b200 import subgraph_sampler
from oracle import oracle as O
from test_sample_assembly import np_edge_rows


class _Ctx:
    def __init__(self, device=0):
        self.device = device

    @classmethod
    def on_torch_stream(cls, device=0):
        return cls(device)

    def sync(self):
        pass

    def edge_rows_host(self, n, src, dst, directed):
        return np_edge_rows(src, dst, n, directed)

    def frontier_distinct(self, cur, cur_slots, prev=()):
        import torch

        return torch.from_numpy(O.np_frontier_distinct(cur.numpy(), cur_slots, [(t.numpy(), s) for t, s in prev]))


class _Graph:
    def __init__(self, rowptr, col, ctx=None):
        self.rowptr, self.col = rowptr, col
        self.n_edges = len(col)
        self.ctx = ctx or _Ctx()

    @classmethod
    def from_edges_host(cls, ctx, n, src, dst, is_graph_directed=True, by_source=False):
        if by_source:  # out-CSR: row u = sorted destinations of u
            return cls(*O.np_build_in_csr(dst, src, n, is_graph_directed), ctx=ctx)
        return cls(*O.np_build_in_csr(src, dst, n, is_graph_directed), ctx=ctx)

    def csr_host(self):
        return self.rowptr, self.col

    def sample_khop_host(self, roots, fanouts, base_seed=42, first_call_no=1):
        return O.c_sample_khop(self.rowptr, self.col, np.asarray(roots, np.int32), list(fanouts), base_seed=base_seed, first_call_no=first_call_no)

    def sample_positives_host(self, srcs, num_pos, base_seed=42, call_no=3):
        got = O.np_sample_positives(self.rowptr, self.col, srcs, num_pos, base_seed=base_seed, call_no=call_no)
        pos = np.full((len(srcs), num_pos), -1, dtype=np.int32)
        cnt = np.zeros(len(srcs), dtype=np.int32)
        for i, u in enumerate(srcs):
            lst = got.get(int(u), [])
            pos[i, :len(lst)] = lst
            cnt[i] = len(lst)
        return pos.reshape(-1), cnt

    def sample_op(self, roots, chain_fanouts, chain_nbr, call_no, base_seed=42, weights=None, method="random_uniform"):
        import torch

        chain = [t.numpy() for t in chain_nbr]
        if method == "random_uniform":
            nbr, cnt = O.np_sample_op((self.rowptr, self.col), roots.numpy(), list(chain_fanouts), chain, call_no, base_seed)
        else:
            nbr, cnt = O.np_sample_op_weighted((self.rowptr, self.col), weights.cpu().numpy(), roots.numpy(), list(chain_fanouts), chain, call_no,
                                               method, base_seed)
        return torch.from_numpy(nbr), torch.from_numpy(cnt)

    def close(self):
        pass


@pytest.fixture(autouse=True)
def oracle_backed_sampler(monkeypatch):
    import torch

    monkeypatch.setattr(subgraph_sampler, "Context", _Ctx)
    monkeypatch.setattr(subgraph_sampler, "Graph", _Graph)
    monkeypatch.setattr(subgraph_sampler, "_roots_to_device", lambda roots, device: torch.from_numpy(np.ascontiguousarray(roots, dtype=np.int32)))
    monkeypatch.setattr(subgraph_sampler, "_floats_to_device", lambda v, device: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)))


def test_node_classification_component(tmp_path):
    G.test_sampler_component_on_reference_fixture(tmp_path)


@pytest.mark.parametrize("directed", [False, True])
def test_link_prediction_component(tmp_path, directed):
    G.test_link_prediction_component_matches_restated_reference(tmp_path, directed)


def test_per_hop_fanouts(tmp_path):
    G.test_per_hop_fanouts_from_sampling_strategy(tmp_path)


def test_user_defined_labels_component(tmp_path):
    G.test_user_defined_labels_component_matches_restated_reference(tmp_path)


@pytest.mark.parametrize("directed", [False, True])
def test_component_sharded_over_ranks(tmp_path, directed):
    G.test_component_sharded_over_ranks_writes_the_same_records(tmp_path, directed)


def test_typed_component_rnn_per_node_type(tmp_path):
    G.test_typed_component_writes_rnn_per_node_type(tmp_path)


def test_typed_component_edge_features_and_isolated_anchors(tmp_path):
    G.test_typed_component_hydrates_edge_features_and_isolated_anchors(tmp_path)


def test_typed_component_on_the_reference_heterogeneous_fixture(tmp_path):
    G.test_typed_component_on_the_reference_heterogeneous_fixture(tmp_path)


def test_typed_component_weighted_sampling_ops(tmp_path):
    G.test_typed_component_weighted_sampling_ops(tmp_path)
