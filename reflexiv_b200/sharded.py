"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the exchanges.

The counting path shards by minimiser bin: every rank bins its own reads into super-k-mer records
laid out shard-major, one all-to-all moves each shard's slice to its owner (this is the step that replaces Spark's
groupBy hash shuffle, ReflexivDataFrameCounter.java:198-200), the owner re-bins and counts locally.  Shard tables are
disjoint, so their concatenation is the global table; no reduction collective is involved.

`exchange_bytes` is backend agnostic (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple


class _DevView:
    """Zero-copy torch view of library-owned device memory (__cuda_array_interface__)."""

    def __init__(self, ptr: int, n_bytes: int):
        self.__cuda_array_interface__ = {"shape": (n_bytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_view(torch, ptr: int, n_bytes: int, device):
    if n_bytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevView(ptr, n_bytes), device=device)


def shard_of_bin(bin_id: int, n_bins_total: int, n_shards: int) -> int:
    """Bins are shard-major: shard s owns bins [s*B/n, (s+1)*B/n)."""
    return bin_id // (n_bins_total // n_shards)


def choose_total_bins(global_instances: int, n_shards: int, target_per_bin: int = 6144, lo: int = 64, hi: int = 1 << 24) -> int:
    """Same rule as the library's single-process choice (rfx_partition.cu: choose_bin_count; 6144 instances per bin for
    k <= 31, 4096 for k > 31), on the GLOBAL instance count, rounded up to a multiple of the shard count so that every
    rank computes the same number.  With a context at hand prefer ``ctx.choose_bins`` (rfx_choose_bins), which asks
    the library itself."""
    nb = max(lo, min(hi, -(-global_instances // target_per_bin)))
    return -(-nb // n_shards) * n_shards


def exchange_bytes(torch, dist, send, send_sizes: Sequence[int], group=None, alloc=None) -> Tuple[object, List[int]]:
    """All-to-all of variable-size byte slices.  `send` is one uint8 tensor holding the slices for rank 0, 1, ...
    back to back.  Returns (recv tensor, recv sizes).  `alloc(n_bytes)` may supply the receive buffer (e.g. a view of
    memory the library owns, so the received records are counted in place)."""
    world = dist.get_world_size(group)
    assert len(send_sizes) == world and sum(send_sizes) == send.numel()
    s = torch.tensor(list(send_sizes), dtype=torch.int64, device=send.device)
    r = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(r, s, group=group)
    recv_sizes = [int(x) for x in r.tolist()]
    recv = alloc(sum(recv_sizes)) if alloc is not None else torch.empty(sum(recv_sizes), dtype=torch.uint8, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_sizes, input_split_sizes=list(send_sizes), group=group)
    return recv, recv_sizes


def gather_varlen(torch, dist, local, group=None):
    """all_gather of tensors of different lengths (same dtype); returns (concatenation in rank order, sizes).
    `local` is 1-D, or [rows, width] with width > 1 (rows vary per rank); sizes are in elements of dim 0.
    One size exchange (a single host read) and one padded all_gather."""
    world = dist.get_world_size(group)
    rows = local.shape[0]
    n = torch.tensor([rows], dtype=torch.int64, device=local.device)
    all_n = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(all_n, n, group=group)
    sizes = [int(x) for x in all_n.tolist()]
    mx = max(sizes) if sizes else 0
    shape = (mx,) if local.dim() == 1 else (mx, local.shape[1])
    if rows == mx:
        pad = local.contiguous()
    else:
        pad = torch.zeros(shape, dtype=local.dtype, device=local.device)
        pad[:rows] = local
    flat = torch.empty(world * pad.numel(), dtype=local.dtype, device=local.device)  # flat buffers: gloo accepts nothing else
    dist.all_gather_into_tensor(flat, pad.reshape(-1), group=group)
    out = flat.reshape((world,) + shape)
    if all(sz == mx for sz in sizes):
        return out.reshape((world * mx,) + shape[1:]), sizes
    return torch.cat([out[r, :sizes[r]] for r in range(world)]), sizes


class _Lap:
    """Optional host-side stopwatch (prof dict): laps are cumulative seconds per label, device-synchronised."""

    def __init__(self, torch, device, prof):
        self.torch, self.device, self.prof = torch, device, prof
        if prof is not None:
            import time
            self.time = time
            self.t = time.perf_counter()

    def __call__(self, label):
        if self.prof is None:
            return
        if getattr(self.device, "type", str(self.device)) != "cpu" and self.torch.cuda.is_available():
            self.torch.cuda.synchronize(self.device)
        now = self.time.perf_counter()
        self.prof[label] = self.prof.get(label, 0.0) + now - self.t
        self.t = now


def _sync(torch, device):
    """The library runs on its own stream and its calls block; torch work that produced its inputs must be done."""
    if getattr(device, "type", str(device)) != "cpu" and torch.cuda.is_available():
        torch.cuda.current_stream(device).synchronize()


def sharded_count(ctx, torch, dist, device, n_bins_total: int, group=None, rebin: bool = False, prof=None) -> dict:
    """Counting across the ranks of `group`; on return `ctx` holds this rank's shard of the global table.
    Every sender's slice is already grouped by bin, so its bin offsets travel with it (a second, tiny all-to-all) and
    the owner counts the received buffer in place; `rebin=True` exercises the general path that re-derives the bins."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lap = _Lap(torch, device, prof)
    ctx.partition(world, n_bins_total)
    lap("count.partition")
    slices = [ctx.shard_records(s) for s in range(world)]
    base = slices[0][0]
    total = sum(n for _, n in slices)
    send = device_view(torch, base, total, device)
    off_ptr, n_off = ctx.shard_bin_offsets(0)
    ctx.begin_shard(rank, world, n_bins_total)
    # receive straight into the context's buffer unless the general (re-binning) path is being exercised
    alloc = None if rebin else (lambda nb: device_view(torch, ctx.rx_buffer(nb), nb, device))
    recv, recv_sizes = exchange_bytes(torch, dist, send, [n for _, n in slices], group, alloc)
    offs_send = device_view(torch, off_ptr, (world * (n_off - 1) + 1) * 8, device).view(torch.int64)
    # shard s needs offsets [s*bps, (s+1)*bps]: bps + 1 values, neighbours share one
    bps = n_off - 1
    idx = (torch.arange(world, device=device).unsqueeze(1) * bps + torch.arange(n_off, device=device).unsqueeze(0)).reshape(-1)
    offs_out = offs_send[idx].contiguous()
    offs_in = torch.empty_like(offs_out)
    dist.all_to_all_single(offs_in, offs_out, group=group)
    _sync(torch, device)
    lap("count.exchange")
    if rebin:
        ctx.load_records_device(recv.data_ptr() if recv.numel() else 0, recv.numel())
    else:
        pos = 0
        for j in range(world):
            ctx.load_segment_device(recv.data_ptr() + pos if recv_sizes[j] else 0, recv_sizes[j], offs_in[j * n_off:(j + 1) * n_off].data_ptr())
            pos += recv_sizes[j]
    del recv
    st = ctx.count()
    lap("count.count")
    return st


def gather_tables(ctx, torch, dist, device, group=None, prof=None) -> dict:
    """Replicates the global (k-mer, count) table on every rank (all_gather of the disjoint shard tables, concatenated
    in rank order) so the graph stages can run; returns the context stats plus `row_ranges`, the rows every rank owns."""
    lap = _Lap(torch, device, prof)
    pk, pc, n, kb = ctx.counts_device()
    keys = device_view(torch, pk, n * kb, device)
    cnts = device_view(torch, pc, n * 4, device)
    all_keys, sizes = gather_varlen(torch, dist, keys, group)
    all_cnts, _ = gather_varlen(torch, dist, cnts, group)
    n_all = sum(sizes) // kb
    _sync(torch, device)
    ctx.load_counts_device(all_keys.data_ptr() if n_all else 0, all_cnts.data_ptr() if n_all else 0, n_all, append=False)
    lap("gather_tables")
    st = ctx.stats()
    lo, ranges = 0, []
    for sz in sizes:
        ranges.append((lo, lo + sz // kb))
        lo += sz // kb
    st["row_ranges"] = ranges
    return st


_UNEVEN_ALL_GATHER = {}


def _bcast_slices(torch, dist, buf, ranges, group=None):
    """Every rank r owns buf[ranges[r][0]:ranges[r][1]]; after the call every rank holds every slice (in place).
    One all_gather over views of `buf` where the backend takes uneven sizes (NCCL: one grouped launch), else one
    broadcast per rank."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    key = dist.get_backend(group)
    if _UNEVEN_ALL_GATHER.get(key, True):
        try:
            views = [buf[lo:hi] for lo, hi in ranges]
            dist.all_gather(views, views[rank].clone(), group=group)
            _UNEVEN_ALL_GATHER[key] = True
            return
        except (RuntimeError, ValueError, NotImplementedError):
            _UNEVEN_ALL_GATHER[key] = False
    for r in range(world):
        lo, hi = ranges[r]
        if hi > lo:
            dist.broadcast(buf[lo:hi], src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def sharded_assemble(ctx, torch, dist, device, row_ranges, group=None, prof=None) -> dict:
    """Graph stages across the ranks of `group` (include/reflexiv_cuda.h: rfx_gs_*): every rank holds the whole table
    (gather_tables) and does the per-node work for its own rows; one byte per node after each fork filter, the
    splitter list of the chain walk and one tuple per chain are what travels.  On return every rank holds the same
    contig set, as after ctx.assemble().  A graph with a closed path falls back to the replicated ctx.assemble()."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    row_lo, row_hi = row_ranges[rank]
    oid_ranges = [(2 * a, 2 * b) for a, b in row_ranges]
    lap = _Lap(torch, device, prof)
    ctx.gs_begin(row_lo, row_hi)
    lap("asm.begin")
    p_alive, n_nodes = ctx.gs_alive()
    alive = device_view(torch, p_alive, n_nodes, device)
    _bcast_slices(torch, dist, alive, oid_ranges, group)
    _sync(torch, device)
    lap("asm.comm")
    ctx.gs_left()
    lap("asm.left")
    _bcast_slices(torch, dist, alive, oid_ranges, group)
    _sync(torch, device)
    lap("asm.comm")
    m, p_node, p_next, p_len = ctx.gs_link()
    lap("asm.link")
    u32 = torch.int32
    trip = torch.empty((m, 3), dtype=u32, device=device)  # one gather for the three arrays
    for j, p in enumerate((p_node, p_next, p_len)):
        if m:
            trip[:, j] = device_view(torch, p, m * 4, device).view(u32)
    g_trip, sizes = gather_varlen(torch, dist, trip, group)
    g_trip = g_trip.t().contiguous()
    g_node, g_next, g_len = g_trip[0], g_trip[1], g_trip[2]
    m_total, my_off = sum(sizes), sum(sizes[:rank])
    _sync(torch, device)
    lap("asm.comm")
    nt, p_t, nh, p_h, has_cycle = ctx.gs_rank(g_node.data_ptr() if m_total else 0, g_next.data_ptr() if m_total else 0, g_len.data_ptr() if m_total else 0,
                                              m_total, my_off)
    lap("asm.rank")
    tails = device_view(torch, p_t, nt * 12, device).view(u32) if nt else torch.empty(0, dtype=u32, device=device)
    heads = device_view(torch, p_h, nh * 8, device).view(u32) if nh else torch.empty(0, dtype=u32, device=device)
    # one gather for both lists and the per-rank scalars: a 6-word header [n_tails, n_heads, has_cycle, n_oriented,
    # n_budget_junctions, n_cycles] (counts < 2^31 per rank), then the tuples
    st0 = ctx.stats()
    H = 6
    hdr_local = torch.tensor([nt, nh, 1 if has_cycle else 0, st0["n_oriented"], st0["n_budget_junctions"], st0["n_cycles"]], dtype=u32, device=device)
    g_blob, bsizes = gather_varlen(torch, dist, torch.cat([hdr_local, tails, heads]), group)
    starts = [sum(bsizes[:r]) for r in range(world)]
    hdr = g_blob[torch.tensor([starts[r] + j for r in range(world) for j in range(H)], device=device)].tolist()
    if any(hdr[H * r + 2] for r in range(world)):
        return ctx.assemble()  # closed paths are opened by the replicated path (rare; none in the BASELINE configs)
    parts_t, parts_h = [], []
    for r in range(world):
        a, b, pos = hdr[H * r], hdr[H * r + 1], starts[r] + H
        parts_t.append(g_blob[pos:pos + 3 * a])
        parts_h.append(g_blob[pos + 3 * a:pos + 3 * a + 2 * b])
    all_tails, all_heads = torch.cat(parts_t), torch.cat(parts_h)
    sums = [sum(hdr[H * r + j] for r in range(world)) for j in (3, 4, 5)]
    nt_all, nh_all = all_tails.numel() // 3, all_heads.numel() // 2
    _sync(torch, device)
    lap("asm.comm")
    p_bases, n_bases = ctx.gs_contigs(all_tails.data_ptr() if nt_all else 0, nt_all, all_heads.data_ptr() if nh_all else 0, nh_all)
    lap("asm.contigs")
    st = ctx.stats()
    if n_bases:
        bases = device_view(torch, p_bases, n_bases, device)
        dist.all_reduce(bases, op=dist.ReduceOp.MAX, group=group)
    adm = torch.tensor([st["n_budget_admissible"]], dtype=torch.int64, device=device)
    dist.all_reduce(adm, group=group)
    _sync(torch, device)
    ctx.gs_finish(sums[0], sums[1], int(adm.item()), sums[2])
    lap("asm.comm")
    return ctx.stats()
