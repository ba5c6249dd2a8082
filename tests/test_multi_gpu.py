"""Multi-rank runs over peer memory (include/reflexiv_cuda.h: rfx_shard_*): reads split across ranks, every rank scans
its reads into slabs over all minimiser bins, the owner of a bin counts it straight out of the peers' slabs, the graph
stages run on the own rows with neighbour probes, chain links and contig bases crossing ranks through peer pointers.
The union of the shard tables, of the fork-filter survivors and of the contigs must equal the single-process oracle
bit for bit.

Two set-ups: ranks as host threads of one process SHARING device 0 (runs on a one-GPU box, which is what the round-end
test box is), and one process per GPU with the arenas mapped through CUDA IPC (needs >= 2 GPUs)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _split(txt: bytes, world: int):
    cuts = [0]
    for r in range(1, world):
        cuts.append(txt.find(b"\n@r", len(txt) * r // world) + 1)
    cuts.append(len(txt))
    return [txt[cuts[r]:cuts[r + 1]] for r in range(world)]


def _rank_body(parts, bins, results):
    def body(rank, ctx):
        ctx.reset()
        if parts[rank]:
            ctx.push_fastq(parts[rank])
        if bins:
            ctx.shard_set_bins(bins)
        st = ctx.count_sharded()
        keys, cnt = ctx.counts()
        st2 = ctx.assemble_sharded()
        sh = ctx.shard_stats()
        ori = ctx.oriented() if not sh["fell_back"] else None
        results[rank] = dict(keys=keys, cnt=cnt, st=st, st2=st2, sh=sh, ori=ori, contigs=ctx.contigs())
    return body


def _check_against_oracle(orc, txt, k, cover, E, min_contig, results):
    from reflexiv_b200.pipeline import keys_to_int
    starts, lens = orc.fastq_reads(txt, orc.FASTQ_RUN)
    c = orc.count_kmers(txt, starts, lens, k, 0, 0, cover, 10_000_000, 1)
    f = orc.fork_filter(c["keys_hi"], c["keys_lo"], c["counts"], k, E)
    a = orc.assemble(f["keys_hi"], f["keys_lo"], f["left"], f["right"], k, min_contig, orc.ASM_CANONICAL)
    ref_ints = [(int(h) << 64) | int(l) for h, l in zip(c["keys_hi"], c["keys_lo"])]
    got = {}
    for res in results:
        for key, n in zip(keys_to_int(res["keys"], k), res["cnt"].tolist()):
            assert key not in got  # shard tables are disjoint
            got[key] = n
    assert sorted(got) == ref_ints
    assert [got[x] for x in ref_ints] == c["counts"].tolist()
    assert sum(r["st"]["n_rows"] for r in results) == len(ref_ints)
    fell_back = any(r["sh"]["fell_back"] for r in results)
    if not fell_back:
        hi = np.concatenate([r["ori"][0] for r in results]); lo = np.concatenate([r["ori"][1] for r in results])
        le = np.concatenate([r["ori"][2] for r in results]); ri = np.concatenate([r["ori"][3] for r in results])
        order = np.lexsort((lo, hi))
        assert np.array_equal(hi[order], f["keys_hi"]) and np.array_equal(lo[order], f["keys_lo"])
        assert np.array_equal(le[order], f["left"]) and np.array_equal(ri[order], f["right"])
    got_c = sorted((s, l, r) for res in results for s, l, r in res["contigs"])
    exp_c = sorted(zip(a["contigs"], a["left"].tolist(), a["right"].tolist()))
    assert got_c == exp_c
    tot = lambda name: sum(r["st2"][name] for r in results)  # noqa: E731
    assert tot("n_oriented") == len(f["left"])
    assert (tot("n_budget_junctions"), tot("n_budget_admissible"), tot("n_cycles")) == (a["n_budget_junctions"], a["n_budget_admissible"], a["n_cycles"])
    assert results[0]["sh"]["n_contigs_global"] == len(a["contigs"])
    return a, fell_back


def _run_threads(R, txt, world, k, cover, E, min_contig, bins=0, arena=1 << 30, device=0):
    from reflexiv_b200 import sharded
    ctxs = [R.ReflexivContext(R.DefaultParam(kmerSize=k, minKmerCoverage=cover, minErrorCoverage=E, minContig=min_contig), device=device) for _ in range(world)]
    try:
        grp = sharded.LocalRanks(ctxs, arena_bytes=arena)
        results = [None] * world
        grp.run(_rank_body(_split(txt, world), bins, results))
        # a second run on the same contexts: arenas, barrier epochs and published blocks are reused
        again = [None] * world
        grp.run(_rank_body(_split(txt, world), bins, again))
        from reflexiv_b200.pipeline import keys_to_int
        for a, b in zip(results, again):  # row order inside a shard table is not fixed (rows leave the counting kernel through an atomic cursor)
            assert dict(zip(keys_to_int(a["keys"], k), a["cnt"].tolist())) == dict(zip(keys_to_int(b["keys"], k), b["cnt"].tolist()))
        assert sorted(x for r in results for x in r["contigs"]) == sorted(x for r in again for x in r["contigs"])
    finally:
        for c in ctxs:
            c.close()
    return results


@pytest.fixture(scope="module")
def R():
    import reflexiv_b200 as R
    return R


def _genome_reads(k, err, seed=0, glen=30_000, pairs=6000, at=False):
    from workload import synth
    g = synth.genome(glen, 400 + k + seed)
    if glen >= 13000:
        g[12000:12800] = g[3000:3800]    # repeat: real forks, budget junctions
    if at:
        g[20000:20040] = np.frombuffer(b"AT" * 20, np.uint8)  # palindromic low-complexity stretch: closed paths
    return bytes(synth.fastq(g, pairs, read_len=150, frag_len=400, error_rate=err, seed_reads=5 + seed, seed_errors=6 + seed))


@pytest.mark.parametrize("world,k,cover,E,err,at", [(2, 31, 2, 8, 0.0, False), (3, 31, 2, 8, 0.01, False), (2, 61, 2, 8, 0.005, False), (4, 21, 1, 8, 0.02, False),
                                                    (2, 31, 2, 0, 0.01, False), (1, 31, 2, 8, 0.01, False), (3, 41, 2, 8, 0.01, True), (2, 24, 1, 8, 0.01, True)])
def test_ranks_sharing_one_device(R, orc, world, k, cover, E, err, at):
    txt = _genome_reads(k, err, at=at)
    results = _run_threads(R, txt, world, k, cover, E, 100)
    a, fell_back = _check_against_oracle(orc, txt, k, cover, E, 100, results)
    assert len(a["contigs"]) >= 2
    assert fell_back == (a["n_cycles"] > 0 and world >= 1) or not at
    if world > 1:
        assert all(r["st"]["n_rows"] > 0 for r in results)
        assert sum(r["sh"]["n_remote_probes"] for r in results) > 0  # some neighbours do live on another rank


def test_docs_golden_contig_over_three_ranks(R, orc, example_text, golden):
    """The reference's documented answer (`reflexiv run -kmer 31 -cover 3` on example/) from a sharded run."""
    import hashlib
    results = _run_threads(R, example_text, 3, 31, 3, 8, 500)
    contigs = [s for r in results for s, _, _ in r["contigs"]]
    assert sorted(len(c) for c in contigs) == [4558, 4558]
    assert sum(c.startswith(golden["documented"]["prefix_1200"]) for c in contigs) == 1
    cs = orc.canonical_contig_set(contigs)
    assert [hashlib.sha256(x.encode()).hexdigest() for x in cs] == golden["oracle"]["contigs_cover3"]["canonical_sha256"]


def test_closed_paths_and_empty_ranks(R, orc):
    """A circular genome has no head: the ranks notice and rank 0 assembles the whole table.  Ranks without reads or rows take part."""
    from workload import synth
    g = synth.genome(2000, 9)
    circ = bytes(g) + bytes(g[:200])
    reads = [circ[i:i + 120] for i in range(0, 2000, 7)] * 2
    txt = "".join(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)).encode()
    results = _run_threads(R, txt, 2, 31, 2, 8, 50)
    a, fell_back = _check_against_oracle(orc, txt, 31, 2, 8, 50, results)
    assert fell_back and a["n_cycles"] >= 1
    # one rank gets no reads at all
    txt2 = _genome_reads(31, 0.0, glen=6000, pairs=600)
    ctxs = [R.ReflexivContext(R.DefaultParam(kmerSize=31, minKmerCoverage=2, minContig=100), device=0) for _ in range(2)]
    from reflexiv_b200 import sharded
    try:
        grp = sharded.LocalRanks(ctxs, arena_bytes=1 << 29)
        results = [None] * 2
        grp.run(_rank_body([txt2, b""], 0, results))
    finally:
        for c in ctxs:
            c.close()
    _check_against_oracle(orc, txt2, 31, 2, 8, 100, results)


def test_dense_forks_across_ranks(R, orc):
    """Budget walks (A9 clauses 3 / 4) that cross ranks: small k on a random genome puts fork winners within reach of each other."""
    from workload import synth
    g = synth.genome(20_000, 311)
    txt = bytes(synth.fastq(g, 4000, read_len=100, frag_len=300, error_rate=0.0, seed_reads=3, seed_errors=4))
    for k in (11, 13):
        results = _run_threads(R, txt, 3, k, 1, 8, k)
        a, _ = _check_against_oracle(orc, txt, k, 1, 8, k, results)
        assert a["n_budget_admissible"] > 0


# ---- one process per GPU, arenas mapped through CUDA IPC ------------------------------------------------------------------
def _worker(rank, world, port, k, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import pickle
    import torch
    import torch.distributed as dist
    import reflexiv_b200 as R
    from reflexiv_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    txt = _genome_reads(k, 0.005, glen=60_000, pairs=12_000)
    ctx = R.ReflexivContext(R.DefaultParam(kmerSize=k, minContig=200), device=rank)
    sharded.connect(ctx, dist, arena_bytes=2 << 30)
    results = [None] * world
    _rank_body(_split(txt, world), 0, results)(rank, ctx)
    with open(os.path.join(out_dir, f"res_{rank}.pkl"), "wb") as f:
        pickle.dump(results[rank], f)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [31, 61, 21])
def test_one_process_per_gpu_over_ipc(tmp_path, orc, k):
    import pickle
    import torch
    import torch.multiprocessing as mp
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    port = 29700 + os.getpid() % 1000 + k
    mp.spawn(_worker, args=(world, port, k, str(tmp_path)), nprocs=world, join=True)
    txt = _genome_reads(k, 0.005, glen=60_000, pairs=12_000)
    results = [pickle.load(open(tmp_path / f"res_{r}.pkl", "rb")) for r in range(world)]
    _check_against_oracle(orc, txt, k, 2, 8, 200, results)


def test_reflexiv_binary_with_gpus_option(tmp_path, golden):
    """`reflexiv run --gpus 3` / `reflexiv counter --gpus 2` (csrc/reflexiv_main.cpp: run_sharded): ranks are host threads of the
    driver, every input file is cut into runs of whole records, one part file per rank.  On a box with fewer devices than
    ranks the ranks share devices (REFLEXIV_ARENA_MB = device memory per rank)."""
    import gzip
    import hashlib
    import subprocess
    from conftest import GOLDEN
    exe = os.path.join(ROOT, "reflexiv_b200", "reflexiv")
    env = dict(os.environ, REFLEXIV_ARENA_MB="768", CUDA_MODULE_LOADING="EAGER", CUDA_DEVICE_MAX_CONNECTIONS="32")
    out = tmp_path / "result"
    r = subprocess.run([exe, "run", "--gpus", "3", "--driver-memory", "3G", "-fastq", os.path.join(GOLDEN, "paired_dat*.fq.gz"), "-outfile", str(out), "-kmer", "31",
                        "-cover", "3"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr + r.stdout
    assert (out / "_SUCCESS").exists()
    recs = "".join((out / f).read_text() for f in sorted(os.listdir(out)) if f.startswith("part-")).strip().split(">")[1:]
    heads = [x.split("\n")[0] for x in recs]
    assert sorted(heads) == ["Contig-4558-(-4,-4)-0", "Contig-4558-(-4,-4)-1"]
    seqs = ["".join(x.strip().split("\n")[1:]) for x in recs]
    assert sum(s.startswith(golden["documented"]["prefix_1200"]) for s in seqs) == 1
    cout = tmp_path / "c"
    r = subprocess.run([exe, "counter", "--gpus", "2", "-fastq", os.path.join(GOLDEN, "paired_dat*.fq.gz"), "-outfile", str(cout), "-kmer", "31", "-cover", "2", "-gzip"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr + r.stdout
    parts = sorted(f for f in os.listdir(cout / "Count_31") if f.startswith("part-"))
    assert len(parts) == 2
    rows = sorted(x for f in parts for x in gzip.open(cout / "Count_31" / f).read().decode().splitlines())
    assert hashlib.sha256(("\n".join(rows) + "\n").encode()).hexdigest() == golden["oracle"]["count_ge2"]["sha256_sorted_csv"]
