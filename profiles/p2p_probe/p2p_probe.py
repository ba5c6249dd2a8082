"""torchrun --nproc-per-node N profiles/p2p_probe/p2p_probe.py : cost of the peer-memory primitives (see p2p_probe.cu)."""
import ctypes as C, os, sys, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
L = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libp2pprobe.so"))
for f in ("p_barrier", "p_bulk", "p_bulk_sum", "p_rate"):
    getattr(L, f).restype = C.c_double
nbytes = 512 << 20
def share(nb):
    p = C.c_void_p(); h = (C.c_ubyte * 64)()
    assert L.p_alloc(rank, C.c_uint64(nb), C.byref(p), h) == 0
    mine = torch.tensor(list(h), dtype=torch.uint8, device="cuda")
    allh = torch.empty(world * 64, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(allh, mine)
    ptrs = []
    for r in range(world):
        if r == rank:
            ptrs.append(p.value); continue
        q = C.c_void_p()
        ph = (C.c_ubyte * 64)(*allh[r * 64:(r + 1) * 64].tolist())
        assert L.p_open(rank, ph, C.byref(q)) == 0, "ipc open failed"
        ptrs.append(q.value)
    return ptrs
data = share(nbytes)
flags = share(4096)
assert L.p_fill(C.c_void_p(data[rank]), C.c_uint64(nbytes), C.c_uint64(0x0101010101010101 * (rank + 1))) == 0
dist.barrier(); torch.cuda.synchronize()
peer = (rank + 1) % world
out = {}
arr = (C.c_void_p * world)(*flags)
out["barrier_us"] = L.p_barrier(arr, rank, world, 200, C.c_ulonglong(1))
dist.barrier()
for name, src in (("local", data[rank]), ("peer", data[peer])):
    for chunk, pieces in ((2048, 8), (16384, 1), (512, 8)):
        out[f"bulk_{name}_{chunk}x{pieces}_GBs"] = L.p_bulk(C.c_void_p(src), C.c_uint64(nbytes), chunk, pieces, 148 * 3)
    dist.barrier()
# correctness of bulk copy from the peer: same checksum as the owner computes locally
mine = L.p_bulk_sum(C.c_void_p(data[rank]), C.c_uint64(nbytes), 2048, 8, 148)
theirs = L.p_bulk_sum(C.c_void_p(data[peer]), C.c_uint64(nbytes), 2048, 8, 148)
t = torch.tensor([mine], dtype=torch.float64, device="cuda"); g = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(g, t)
out["bulk_peer_checksum_ok"] = bool(abs(g[peer].item() - theirs) < 0.5)
dist.barrier()
cnt = 32 << 20
scratch = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for name, ptr in (("local", data[rank]), ("peer", data[peer])):
    out[f"rand_read8_{name}_Gops"] = cnt / L.p_rate(0, C.c_void_p(ptr), C.c_uint64(nbytes), C.c_uint64(cnt), None) / 1e6
    out[f"rand_store1_{name}_Gops"] = cnt / L.p_rate(1, C.c_void_p(ptr), C.c_uint64(nbytes), C.c_uint64(cnt), None) / 1e6
    out[f"run10_store1_{name}_Gbytes"] = 10 * (cnt // 8) / L.p_rate(2, C.c_void_p(ptr), C.c_uint64(nbytes), C.c_uint64(cnt // 8), None) / 1e6
    out[f"rand_store8_{name}_Gops"] = cnt / L.p_rate(3, C.c_void_p(ptr), C.c_uint64(nbytes), C.c_uint64(cnt), None) / 1e6
    out[f"stream_copy_from_{name}_GBs"] = nbytes / L.p_rate(4, C.c_void_p(ptr), C.c_uint64(nbytes), C.c_uint64(0), C.c_void_p(scratch.data_ptr())) / 1e6
    dist.barrier()
print(f"rank {rank} world {world}: " + ", ".join(f"{k}={v:.3f}" if isinstance(v, float) else f"{k}={v}" for k, v in out.items()), flush=True)
dist.barrier()
dist.destroy_process_group()
