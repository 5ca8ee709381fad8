"""N > 1 host logic on CPU: world_size-2 gloo processes shard the roots into disjoint contiguous
ranges that cover every node exactly once, and agree on the max-over-ranks time."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gigl_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_nodes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.root_range(n_nodes, rank, world)
    ranges = [None] * world
    dist.all_gather_object(ranges, (lo, hi))
    batches = sharding.root_batches(n_nodes, rank, world, 7, 5)
    t = sharding.max_over_ranks(1.0 + rank)
    q.put((rank, ranges, [b.tolist() for b in batches], t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_root_sharding_and_timing_reduction():
    world, n_nodes = 2, 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_nodes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = res[0][1]
    assert ranges == res[1][1] and ranges[0][0] == 0 and ranges[-1][1] == n_nodes and ranges[0][1] == ranges[1][0]
    for rank, _, batches, t in res:
        lo, hi = ranges[rank]
        flat = np.array(batches).ravel()
        assert flat.min() >= lo and flat.max() < hi
        assert t == 2.0  # the slowest rank's time everywhere
    # single process: ranges tile [0, N) for any world size, batches walk the range in id order
    for w in (1, 3, 8):
        cover = np.concatenate([np.arange(*sharding.root_range(1000, r, w)) for r in range(w)])
        assert np.array_equal(cover, np.arange(1000))
    b = sharding.root_batches(10, 0, 1, 4, 4)
    assert [x.tolist() for x in b] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 0, 1], [2, 3, 4, 5]]


def _fd_worker(rank, world, tag, q):
    # every rank owns a pipe; the peers receive its WRITE end through exchange_fds and write their rank into it
    r_fd, w_fd = os.pipe()
    got = sharding.exchange_fds(w_fd, rank, world, tag, timeout=60)
    assert sorted(got) == [p for p in range(world) if p != rank]
    for peer, fd in got.items():
        os.write(fd, bytes([rank]))
        os.close(fd)
    seen = sorted(os.read(r_fd, 1)[0] for _ in range(world - 1))
    q.put((rank, seen))


def test_fd_exchange_between_processes_and_shard_arithmetic():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tag = f"test{os.getpid()}"
    procs = [ctx.Process(target=_fd_worker, args=(r, world, tag, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: [1, 2], 1: [0, 2], 2: [0, 1]}
    # shards: multiples of the granule, together covering every node
    assert sharding.shard_rows(2_449_029, 8, 131072) == 393216 and 8 * 393216 >= 2_449_029
    assert sharding.shard_rows(1000, 1, 4096) == 4096 and sharding.shard_rows(8192, 2, 4096) == 4096
    assert sharding.exchange_fds(0, 0, 1, tag) == {}
