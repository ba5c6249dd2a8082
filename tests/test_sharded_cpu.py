"""World-size-2 gloo tests (CPU) of the sharding logic: bins laid out shard-major, one all-to-all of the record
slices, owners count what they receive, the union of the shard tables is the global table.  The records are produced
by the host harness from the same rfx_core.h code the kernels use; on GPUs the same plumbing runs over NCCL
(reflexiv_b200/sharded.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K, M = 31, 11


def _worker(rank, world, port, n_bins_total, result_dir):
    import ctypes as C
    from conftest import make_reads
    from oracle import orc
    from reflexiv_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = C.CDLL(os.path.join(ROOT, "tests", "hostemu", "libhostemu.so"))
    P, I64, U32, I = C.c_void_p, C.c_int64, C.c_uint32, C.c_int
    L.emu_pack_reads.restype = I64; L.emu_pack_reads.argtypes = [P, P, P, I64, I, I, I, P, P, P]
    L.emu_partition.restype = I64; L.emu_partition.argtypes = [P, I64, P, P, I64, I, I, U32]
    L.emu_partition_fetch.argtypes = [P, P]
    L.emu_count_records.restype = I64; L.emu_count_records.argtypes = [P, P, I64, I, I, U32, P]
    L.emu_count_fetch.argtypes = [P, P, P]
    # every rank has its own reads of one shared genome
    txt = make_reads(77, 20_000, 1500, read_len=100, err=0.01)
    half = len(txt) // 2
    cut = bytes(txt).find(b"\n@r", half) + 1
    mine = bytes(txt)[:cut] if rank == 0 else bytes(txt)[cut:]
    starts, lens = orc.fastq_reads(mine, orc.FASTQ_RUN)
    a = np.frombuffer(mine, np.uint8)
    elen = np.zeros(len(starts), np.uint32); woff = np.zeros(len(starts), np.uint64)
    nw = L.emu_pack_reads(a.ctypes.data, starts.ctypes.data, lens.ctypes.data, len(starts), K, 0, 0, elen.ctypes.data, woff.ctypes.data, None)
    words = np.zeros(nw + 8, np.uint64)
    L.emu_pack_reads(a.ctypes.data, starts.ctypes.data, lens.ctypes.data, len(starts), K, 0, 0, elen.ctypes.data, woff.ctypes.data, words.ctypes.data)
    n = L.emu_partition(words.ctypes.data, len(words), elen.ctypes.data, woff.ctypes.data, len(elen), K, M, n_bins_total)
    recs = np.zeros((n, 2), np.uint64); bins = np.zeros(n, np.uint32)
    L.emu_partition_fetch(recs.ctypes.data, bins.ctypes.data)
    # shard-major layout: sort records by bin, slice per shard
    order = np.argsort(bins, kind="stable")
    recs, bins = recs[order], bins[order]
    shard = np.array([sharded.shard_of_bin(int(b), n_bins_total, world) for b in bins])
    sizes = [int((shard == s).sum()) * 16 for s in range(world)]
    send = torch.from_numpy(recs.view(np.uint8).reshape(-1).copy())
    recv, recv_sizes = sharded.exchange_bytes(torch, dist, send, sizes)
    assert sum(recv_sizes) == recv.numel()
    got = recv.numpy().view(np.uint64).reshape(-1, 2).copy()
    # every received record must belong to one of my bins; count them
    bad = C.c_int64(0)
    d = L.emu_count_records(got.ctypes.data, None, len(got), K, M, n_bins_total, C.addressof(bad))
    hi = np.zeros(d, np.uint64); lo = np.zeros(d, np.uint64); cnt = np.zeros(d, np.uint32)
    L.emu_count_fetch(hi.ctypes.data, lo.ctypes.data, cnt.ctypes.data)
    # gather the (disjoint) shard tables everywhere, as the GPU path does before the graph stages
    keys_all, sz = sharded.gather_varlen(torch, dist, torch.from_numpy(lo.view(np.int64)))
    cnts_all, _ = sharded.gather_varlen(torch, dist, torch.from_numpy(cnt.astype(np.int64)))
    np.save(os.path.join(result_dir, f"keys_{rank}.npy"), keys_all.numpy().view(np.uint64))
    np.save(os.path.join(result_dir, f"cnts_{rank}.npy"), cnts_all.numpy())
    np.save(os.path.join(result_dir, f"local_{rank}.npy"), lo)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharded_count_matches_oracle(tmp_path, orc, hostemu):
    from conftest import make_reads
    from reflexiv_b200 import sharded
    world = 2
    n_bins_total = sharded.choose_total_bins(2 * 1500 * 70, world, target_per_bin=4096)
    assert n_bins_total % world == 0
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, n_bins_total, str(tmp_path)), nprocs=world, join=True)
    txt = make_reads(77, 20_000, 1500, read_len=100, err=0.01)
    ref = orc.count_kmers(txt, *orc.fastq_reads(txt, orc.FASTQ_RUN), K)
    for rank in range(world):
        keys = np.load(tmp_path / f"keys_{rank}.npy")
        cnts = np.load(tmp_path / f"cnts_{rank}.npy")
        order = np.argsort(keys)
        assert np.array_equal(keys[order], ref["keys_lo"])        # union of the shards = global table, no duplicates
        assert np.array_equal(cnts[order].astype(np.uint32), ref["counts"])
    l0, l1 = np.load(tmp_path / "local_0.npy"), np.load(tmp_path / "local_1.npy")
    assert len(np.intersect1d(l0, l1)) == 0 and len(l0) > 0 and len(l1) > 0


def test_bin_geometry_helpers():
    from reflexiv_b200 import sharded
    for world in (1, 2, 4, 8):
        nb = sharded.choose_total_bins(368_000_160 * world, world)
        assert nb % world == 0 and nb >= 64
        per = nb // world
        assert [sharded.shard_of_bin(b, nb, world) for b in (0, per - 1, per if world > 1 else 0, nb - 1)][-1] == world - 1
    assert sharded.choose_total_bins(10, 8) == 64


def _plumbing_worker(rank, world, port):
    from reflexiv_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # gather_varlen: ragged 1-D and 2-D inputs come back concatenated in rank order
    rows = 3 + 2 * rank
    one = torch.arange(rows, dtype=torch.int64) + 100 * rank
    g, sizes = sharded.gather_varlen(torch, dist, one)
    assert sizes == [3 + 2 * r for r in range(world)]
    assert g.tolist() == [100 * r + i for r in range(world) for i in range(3 + 2 * r)]
    two = (torch.arange(rows * 3, dtype=torch.int32) + 1000 * rank).reshape(rows, 3)
    g2, sizes2 = sharded.gather_varlen(torch, dist, two)
    assert sizes2 == sizes and g2.shape == (sum(sizes), 3)
    assert g2[sizes[0]:sizes[0] + 2].tolist() == [[1000, 1001, 1002], [1003, 1004, 1005]]
    empty, esz = sharded.gather_varlen(torch, dist, torch.empty(0, dtype=torch.int32))
    assert empty.numel() == 0 and esz == [0] * world
    # _bcast_slices: every rank ends up with every owner's slice, whatever path the backend allows
    ranges = [(0, 5), (5, 12)]
    buf = torch.zeros(12, dtype=torch.uint8)
    lo, hi = ranges[rank]
    buf[lo:hi] = rank + 1
    sharded._bcast_slices(torch, dist, buf, ranges)
    assert buf.tolist() == [1] * 5 + [2] * 7
    dist.barrier()
    dist.destroy_process_group()


def test_collective_helpers_on_gloo():
    port = 29800 + os.getpid() % 1000
    mp.spawn(_plumbing_worker, args=(2, port), nprocs=2, join=True)
