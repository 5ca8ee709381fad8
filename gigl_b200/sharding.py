"""Root sharding for N > 1 GPUs: sampling units (roots) are independent, so rank r of P
owns the contiguous id range [r*N/P, (r+1)*N/P) - the rule the reference uses for its loaders
(python/gigl/distributed/distributed_neighborloader.py:195-216) - and no data-path collective is needed.
The only cross-rank step of a measurement is the max-over-ranks of the device time."""
from __future__ import annotations

import numpy as np


def root_range(n_nodes: int, rank: int, world: int):
    return rank * n_nodes // world, (rank + 1) * n_nodes // world


def root_batches(n_nodes: int, rank: int, world: int, batch: int, n_steps: int, start_step: int = 0):
    """Step s of rank r = the next `batch` ids of the rank's range, in id order, wrapping inside the range."""
    lo, hi = root_range(n_nodes, rank, world)
    span = max(hi - lo, 1)
    return [(lo + (np.arange(batch, dtype=np.int64) + s * batch) % span).astype(np.int32)
            for s in range(start_step, start_step + n_steps)]


def max_over_ranks(value: float, device=None) -> float:
    """The contract's timing reduction: every rank reports the slowest rank's time."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------------------------
# Node features sharded over the GPUs of the box, one flat array on every GPU (csrc/shared_table.cu)
# ----------------------------------------------------------------------------------------------
def shard_rows(n_nodes: int, world: int, granule: int) -> int:
    """Rows per shard: ceil(n_nodes / world) rounded up to the mapping granule (shard k = rows [k*rps, (k+1)*rps))."""
    per = -(-n_nodes // world)
    return -(-per // granule) * granule


def _uds_dir(tag: str) -> str:
    """Directory of the job's fd sockets: private to this user (0700, ownership checked), one per job tag."""
    import os
    import stat
    import tempfile

    d = os.path.join(tempfile.gettempdir(), f"gigl_b200_{os.getuid()}_{tag}")
    try:
        os.mkdir(d, 0o700)
    except FileExistsError:
        pass
    st = os.lstat(d)
    if not stat.S_ISDIR(st.st_mode) or st.st_uid != os.getuid() or (st.st_mode & 0o077):
        raise PermissionError(f"{d} is not a private directory of this user")
    return d


def _peer_uid(conn) -> int:
    import socket
    import struct

    cred = conn.getsockopt(socket.SOL_SOCKET, socket.SO_PEERCRED, struct.calcsize("3i"))
    return struct.unpack("3i", cred)[1]


def exchange_fds(my_fd: int, rank: int, world: int, tag: str, timeout: float = 120.0):
    """Every rank hands a duplicate of ``my_fd`` to every peer over Unix sockets (SCM_RIGHTS) and gets theirs:
    returns {peer_rank: fd}.  All processes must be on one host (one box = one NVSwitch domain).  The fd maps this GPU's
    feature shard read/write, so the sockets live in a 0700 directory of this user and a server only answers a peer
    whose SO_PEERCRED uid is its own."""
    import os
    import socket
    import threading
    import time

    if world == 1:
        return {}
    d = _uds_dir(tag)
    path = os.path.join(d, f"{rank}.sock")
    try:
        os.unlink(path)
    except FileNotFoundError:
        pass
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(world)
    srv.settimeout(timeout)

    def serve():
        served = 0
        while served < world - 1:
            conn, _ = srv.accept()
            with conn:
                if _peer_uid(conn) != os.getuid():
                    continue
                conn.recv(4)
                socket.send_fds(conn, [b"fd"], [my_fd])
                served += 1

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    got = {}
    for peer in range(world):
        if peer == rank:
            continue
        deadline = time.time() + timeout
        while True:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            try:
                c.connect(os.path.join(d, f"{peer}.sock"))
                break
            except (ConnectionRefusedError, FileNotFoundError):
                c.close()
                if time.time() > deadline:
                    raise TimeoutError(f"rank {rank}: peer {peer} never opened its fd socket")
                time.sleep(0.05)
        with c:
            c.sendall(b"gimm")
            _, fds, _, _ = socket.recv_fds(c, 16, 1)
            got[peer] = fds[0]
    th.join(timeout)
    srv.close()
    try:
        os.unlink(path)
    except FileNotFoundError:
        pass
    return got


def hot_rows(graph, x, fraction: float):
    """The hot set of a sharded feature table: the ceil(fraction * N) vertices that occur most often as a SOURCE of the
    graph's edges (a vertex is drawn into a batch through the in-edge rows it sits in, so its out-degree is its weight;
    on an undirected graph that is its degree).  Returns (slot int32 [N]: vertex -> row of the copy or -1, table
    [n_hot, pitch] with the rows copied out of `x` - peer rows cross NVLink once, here)."""
    import torch

    n = graph.n_nodes
    n_hot = min(n, int(-(-fraction * n // 1))) if fraction > 0 else 0
    dev = x.device
    slot = torch.full((n,), -1, dtype=torch.int32, device=dev)
    if n_hot == 0:
        return slot, torch.empty((0, x.shape[1]), dtype=torch.float32, device=dev)
    _, col = graph.csr_tensors()
    deg = torch.zeros(n, dtype=torch.int64, device=dev)
    step = 1 << 27
    for s0 in range(0, col.numel(), step):  # out-degree, in chunks (a 1e9-edge column list is 8 GB as int64)
        deg += torch.bincount(col[s0:s0 + step].long(), minlength=n)
    ids = torch.topk(deg, n_hot, sorted=False).indices.sort().values
    del deg
    slot[ids] = torch.arange(n_hot, dtype=torch.int32, device=dev)
    pitch = -(-x.shape[1] // 32) * 32
    table = torch.zeros((n_hot, pitch), dtype=torch.float32, device=dev)
    for r0 in range(0, n_hot, 1 << 20):
        r1 = min(n_hot, r0 + (1 << 20))
        table[r0:r1, : x.shape[1]] = x[ids[r0:r1]]
    return slot, table[:, : x.shape[1]]


class ShardedFeatureTable:
    """The [n_nodes, F] fp32 feature table with shard ``rank`` resident on this GPU and every other shard mapped from
    its owner's memory, as ONE flat CUDA array (``.table``, a torch view usable by ``Graph.set_features`` /
    ``Batch.sage_forward``).  ``.local`` is this rank's slice (rows [rank * rows_per_shard, ...)) to fill."""

    def __init__(self, ctx, n_nodes: int, F: int, rank: int, world: int, tag: str = "0"):
        import ctypes as C
        import os

        import torch

        from ._capi import check
        from .engine import _tensor_from_ptr

        L = ctx._L
        self.ctx, self.n_nodes, self.F, self.rank, self.world = ctx, n_nodes, F, rank, world
        g = C.c_int64()
        check(L.gigl_shared_table_row_granule(ctx.handle, F, C.byref(g)), ctx.handle)
        self.rows_per_shard = shard_rows(n_nodes, world, g.value)
        h, fd = C.c_void_p(), C.c_int32(-1)
        check(L.gigl_shared_table_create(ctx.handle, world, rank, self.rows_per_shard, F, C.byref(h), C.byref(fd)), ctx.handle)
        self.handle = h
        peers = exchange_fds(fd.value, rank, world, tag)
        for peer, pfd in sorted(peers.items()):
            try:
                check(L.gigl_shared_table_attach(h, peer, pfd), ctx.handle)
            finally:
                os.close(pfd)  # the driver keeps its own reference to the allocation
        base, mine, total = C.c_void_p(), C.c_void_p(), C.c_int64()
        L.gigl_shared_table_ptrs(h, C.byref(base), C.byref(mine), C.byref(total))
        dev = torch.device("cuda", ctx.device)
        self.table = _tensor_from_ptr(base.value, (total.value, F), torch.float32, dev, owner=self)
        self.local = _tensor_from_ptr(mine.value, (self.rows_per_shard, F), torch.float32, dev, owner=self)
        self.row_lo = rank * self.rows_per_shard
        self.row_hi = max(self.row_lo, min(n_nodes, (rank + 1) * self.rows_per_shard))

    def close(self):
        if getattr(self, "handle", None):
            self.table = self.local = None
            self.ctx._L.gigl_shared_table_destroy(self.handle)
            self.handle = None
