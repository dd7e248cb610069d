"""world_size-2 gloo test of the multi-partition protocol bench.py uses at N > 1 GPUs:
each rank owns one partition, every query goes to every partition, per-rank top-k are
all-gathered and merged in the router's order (internal/client/client.go:1530-1609).  On CPU the
per-partition search and the merge are the oracle's; the property checked is the protocol's:
merged(partition results) == top-k of the union, ids = (partition << 32) | local id."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from vearch_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, nq, k = 3000, 16, 12, 10
        part = synth.sift_like(n, d, seed=100 + rank)  # this rank's partition
        xq = synth.sift_like(nq, d, seed=7)            # the same queries on every rank
        dis, ids = orc.flat_search(part, xq, k, orc.METRIC_L2)
        gd = [torch.empty((nq, k), dtype=torch.float32) for _ in range(world)]
        gi = [torch.empty((nq, k), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gd, torch.from_numpy(dis))
        dist.all_gather(gi, torch.from_numpy(ids))
        md, mi = orc.merge_partitions(torch.stack(gd).numpy(), torch.stack(gi).numpy(), orc.METRIC_L2)
        # every rank must hold the same merged answer
        t = torch.from_numpy(mi.copy())
        ref = t.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(t, ref)
        if rank == 0:
            np.savez(os.path.join(out_dir, "merged.npz"), md=md, mi=mi)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_partitions_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "merged.npz"))
    parts = [synth.sift_like(3000, 16, seed=100 + r) for r in range(world)]
    xq = synth.sift_like(12, 16, seed=7)
    union = np.concatenate(parts)
    dis, ids = orc.flat_search(union, xq, 10, orc.METRIC_L2)
    assert np.array_equal(got["md"], dis)  # same scores as one search over the union
    part_of = got["mi"] >> 32
    local = got["mi"] & 0xFFFFFFFF
    glob = part_of * 3000 + local
    for q in range(xq.shape[0]):  # same documents, up to the router's tie order between partitions
        assert sorted(glob[q]) == sorted(ids[q]) or len(set(dis[q])) < 10


# ---- the key protocol bench.py uses since round 2: ONE all-gather of 64-bit result keys -------------------------------
def _f2ord(f):
    """order-preserving float32 -> uint32 (csrc/common.cuh f2ord): smaller key == smaller L2 score"""
    b = np.ascontiguousarray(f, np.float32).view(np.uint32).astype(np.uint64)
    return np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000)


def _ord2f(o):
    o = o.astype(np.uint64)
    b = np.where(o & 0x80000000, o & 0x7FFFFFFF, (~o) & 0xFFFFFFFF).astype(np.uint32)
    return b.view(np.float32)


def _key_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, nq, k = 3000, 16, 12, 10
        part = synth.sift_like(n, d, seed=100 + rank)
        xq = synth.sift_like(nq, d, seed=7)
        dis, ids = orc.flat_search(part, xq, k, orc.METRIC_L2)
        # what gb_index_search_device_keys hands out: (order-preserving score bits << 32) | local id, sorted ascending
        keys = ((_f2ord(dis) << np.uint64(32)) | ids.astype(np.uint64)).astype(np.uint64)
        mine = torch.from_numpy(keys.view(np.int64).copy())
        # the one collective of a step, 8 B per (query, result): all_gather_into_tensor on NCCL; gloo wants the flat form
        gathered = torch.empty((world * nq, k), dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, mine)
        g = gathered.numpy().view(np.uint64).reshape(world, nq, k)
        # merge_partitions_kernel<FROM_KEYS>: order by (score bits, later partition first, rank inside the partition)
        md = np.empty((nq, k), np.float32)
        mi = np.empty((nq, k), np.int64)
        for q in range(nq):
            rows = sorted(((int(g[p, q, j] >> np.uint64(32)), world - 1 - p, j, p, int(g[p, q, j] & np.uint64(0xFFFFFFFF)))
                           for p in range(world) for j in range(k)))[:k]
            md[q] = _ord2f(np.array([r[0] for r in rows], np.uint64))
            mi[q] = [(r[3] << 32) | r[4] for r in rows]
        # the two-array protocol of round 1 (scores + ids, two collectives) must give the same answer
        gd = [torch.empty((nq, k), dtype=torch.float32) for _ in range(world)]
        gi = [torch.empty((nq, k), dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gd, torch.from_numpy(dis))
        dist.all_gather(gi, torch.from_numpy(ids))
        od, oi = orc.merge_partitions(torch.stack(gd).numpy(), torch.stack(gi).numpy(), orc.METRIC_L2)
        assert np.array_equal(md, od) and np.array_equal(mi, oi)
        if rank == 0:
            np.savez(os.path.join(out_dir, "merged_keys.npz"), md=md, mi=mi)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_partitions_gloo_key_protocol(tmp_path):
    world = 2
    mp.spawn(_key_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "merged_keys.npz"))
    parts = [synth.sift_like(3000, 16, seed=100 + r) for r in range(world)]
    xq = synth.sift_like(12, 16, seed=7)
    dis, _ = orc.flat_search(np.concatenate(parts), xq, 10, orc.METRIC_L2)
    assert np.array_equal(got["md"], dis)
