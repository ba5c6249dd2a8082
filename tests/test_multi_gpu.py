"""NCCL path on real GPUs (needs >= 2): reads split across ranks, super-k-mer records exchanged with one all-to-all,
shards counted locally, tables all-gathered, graph stages sharded by row owner (fork filters, links, chain walk and
base gather on the own rows; alive bytes, splitter list and chain tuples exchanged) or, for comparison, replicated.
The union of the shard tables and the contigs must equal the single-process oracle bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, k, rebin, graph, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import reflexiv_b200 as R
    from reflexiv_b200 import sharded
    from conftest import make_reads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    txt = bytes(make_reads(31, 60_000, 12_000, read_len=150, err=0.005, frag=400))
    # rank r takes the r-th slice of the records (cut on record boundaries)
    cuts = [0]
    for r in range(1, world):
        cuts.append(txt.find(b"\n@r", len(txt) * r // world) + 1)
    cuts.append(len(txt))
    mine = txt[cuts[rank]:cuts[rank + 1]]
    ctx = R.ReflexivContext(R.DefaultParam(kmerSize=k, minContig=200), device=rank)
    ctx.push_fastq(mine)
    inst = torch.tensor([ctx.stats()["n_instances"]], dtype=torch.int64, device=device)
    dist.all_reduce(inst)
    n_bins_total = sharded.choose_total_bins(int(inst.item()), world, 4096)
    st = sharded.sharded_count(ctx, torch, dist, device, n_bins_total, rebin=rebin)
    keys, cnt = ctx.counts()
    np.save(os.path.join(out_dir, f"keys_{rank}.npy"), keys)
    np.save(os.path.join(out_dir, f"cnt_{rank}.npy"), cnt)
    gst = sharded.gather_tables(ctx, torch, dist, device)
    if graph == "sharded":
        st2 = sharded.sharded_assemble(ctx, torch, dist, device, gst["row_ranges"])
    else:
        st2 = ctx.assemble()
    np.save(os.path.join(out_dir, f"asm_stats_{rank}.npy"), np.array([st2["n_oriented"], st2["n_contigs"], st2["n_contig_bases"], st2["n_budget_junctions"],
                                                                      st2["n_budget_admissible"], st2["n_cycles"]], dtype=np.int64))
    contigs = sorted(c for c, _, _ in ctx.contigs())
    with open(os.path.join(out_dir, f"contigs_{rank}.txt"), "w") as f:
        f.write("\n".join(contigs))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k,rebin,graph", [(31, False, "sharded"), (61, False, "sharded"), (31, True, "replicated"), (21, False, "sharded")])
def test_sharded_count_and_assembly_over_nccl(tmp_path, orc, k, rebin, graph):
    import torch
    import torch.multiprocessing as mp
    from conftest import make_reads
    from reflexiv_b200.pipeline import keys_to_int
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    port = 29700 + os.getpid() % 1000 + k + (7 if rebin else 0)
    mp.spawn(_worker, args=(world, port, k, rebin, graph, str(tmp_path)), nprocs=world, join=True)
    txt = bytes(make_reads(31, 60_000, 12_000, read_len=150, err=0.005, frag=400))
    ref = orc.run_pipeline(txt, k=k, cover=2, min_contig=200)
    c = ref["counts"]
    ref_ints = [(int(h) << 64) | int(l) for h, l in zip(c["keys_hi"], c["keys_lo"])]
    got = {}
    for r in range(world):
        keys = np.load(tmp_path / f"keys_{r}.npy")
        cnt = np.load(tmp_path / f"cnt_{r}.npy")
        assert len(keys) > 0
        for key, n in zip(keys_to_int(keys, k), cnt.tolist()):
            assert key not in got          # shard tables are disjoint
            got[key] = n
    assert sorted(got) == ref_ints
    assert [got[x] for x in ref_ints] == c["counts"].tolist()
    expect = "\n".join(sorted(ref["asm"]["contigs"]))
    for r in range(world):
        assert (tmp_path / f"contigs_{r}.txt").read_text() == expect
    stats = [np.load(tmp_path / f"asm_stats_{r}.npy").tolist() for r in range(world)]
    assert all(s == stats[0] for s in stats)                                  # every rank reports the global numbers
    a = ref["asm"]
    assert stats[0][1] == len(a["contigs"]) and stats[0][2] == sum(len(x) for x in a["contigs"])
