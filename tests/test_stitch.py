"""`reflexiv run -stitch` (ReflexivDSMain.java:585-672): oracle against a literal Python transcription of the
Java loops, the CUDA kernels run on the host thread by thread against the oracle (no GPU needed), and the library through
the C ABI and both drivers against the oracle (GPU).

PARITY UNPINNED: the reference holds no vector for this branch; see oracle/stitch_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- inputs ---------------------------------------------------------------------------------------------------------
def fq(reads):
    return "".join(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n" for i, r in enumerate(reads)).encode()


def gap_scenario(seed, glen, L, gaps, circular=False, err=0.0, extra=()):
    """Reads at high coverage everywhere except over `gaps`, each of which only two reads (one per strand) bridge."""
    rng = np.random.default_rng(seed)
    G = "".join(rng.choice(list("ACGT"), glen))
    GG = G + G[:L] if circular else G
    reads, prev, segs = [], 0, []
    for a, b in gaps:
        segs.append((prev, a))
        prev = b
    segs.append((prev, len(GG)))
    for a, b in segs:
        reads += [GG[s:s + L] for s in range(a, b - L + 1, 3)]
    for a, b in gaps:
        for j in range(2):
            r = GG[max(0, a - L // 2 + 5 * j):b + L // 2 + 3 * j]
            if err:
                r = "".join(c if rng.random() > err else "ACGT"[rng.integers(4)] for c in r)
            reads.append(r if j == 0 else orc.revcomp_str(r))
    return G, fq(reads + list(extra))


def random_scenarios(n):
    for s in range(n):
        rng = np.random.default_rng(100 + s)
        k = int(rng.choice([11, 13, 15, 21, 31]))
        glen = int(rng.integers(1200, 4000))
        cand = np.arange(300, glen - 300, 150)
        cuts = sorted(rng.choice(cand, size=min(int(rng.integers(1, 5)), len(cand)), replace=False).tolist())
        gaps = [(c, c + int(rng.integers(k + 5, 70))) for c in cuts]
        yield k, gap_scenario(100 + s, glen, 90, gaps, circular=bool(s % 2), err=0.01 if s % 3 == 0 else 0.0)[1]


def natural_scenario(glen, coverage, seed=7, err=0.0):
    """Uniformly sampled paired reads at LOW coverage: the k-mer coverage drops below -cover here and there, which is what
    the stitch branch is for."""
    from workload import synth
    g = synth.genome(glen, seed=seed)
    return bytes(synth.fastq(g, synth.n_pairs_for(glen, coverage, 150), read_len=150, frag_len=400, error_rate=err))


def assembled(txt, k, cover=3):
    asm = orc.run_pipeline(txt, k=k, cover=cover, min_contig=0)["asm"]
    return asm["contigs"], asm["left"].tolist(), asm["right"].tolist()


def triples(res):
    return sorted(zip(res["contigs"], np.asarray(res["left"]).tolist(), np.asarray(res["right"]).tolist()))


# ---- a literal transcription of the Java (strings and dicts, the control flow of the reference line by line) ----------
def _nv(c):  # nucleotideValue, ReflexivDSMain.java:1597-1609
    return "ACGT"[0 if c == "A" else 1 if c == "C" else 2 if c == "G" else 3]


def _comp(a):  # complementary, :1527-1539
    return "T" if a in "Aa" else "A" if a in "TtUu" else "G" if a in "Cc" else "C" if a in "Gg" else "N"


def py_stitch(contigs, left, right, reads, k, min_contig):
    sk = k - 1
    # DSLowCoverageSubKmerExtraction :1211-1268; records arrive sorted by their first k-mer (CANONICAL ORDER), Hashtable.put
    table = {}
    order = sorted(range(len(contigs)), key=lambda c: contigs[c][:k])
    for c in order:
        s = contigs[c]
        if len(s) < 61:
            continue
        if -5 <= left[c] < 0:
            table[s[:sk]] = (c, 1)
        if -5 <= right[c] < 0:
            table[s[-sk:]] = (c, 0)
    frags = []
    for rd in reads:  # DSLowCoverageReadDetection.call :1463-1543
        if len(rd) - sk <= 1:
            continue
        for strand in (0, 1):
            x = "".join(_nv(ch) for ch in (rd if strand == 0 else "".join(_comp(ch) for ch in reversed(rd))))
            probed, pl, pr = -1, -1, -1
            for i in range(len(x)):
                if i < sk - 1:
                    continue
                hit = table.get(x[i - sk + 1:i + 1])
                if hit is None:
                    continue
                marker, direction = hit
                if direction == 0:
                    if pl == -1:
                        probed, pl = marker, i
                else:
                    if probed == marker:
                        continue
                    pr = i
            if pl >= 0 and pr >= 0 and pl < pr:
                f = x[pl - sk + 1:pr + 1]
                frags.append((len(f), f, table[f[:sk]][0], table[f[-sk:]][0]))
    nxt, prv = {}, {}
    for f in frags:  # one record per first (k-1)-mer (DSFilterRepeatLowCoverageFragment), CANONICAL ORDER: the smallest
        if f[2] not in nxt or f[:2] < nxt[f[2]][:2]:
            nxt[f[2]] = f
    for f in nxt.values():  # of the survivors that end on one contig the smallest joins it
        if f[3] not in prv or f[:2] < prv[f[3]][:2]:
            prv[f[3]] = f
    out, done = [], set()

    def walk(c):
        seq, r, cur = contigs[c], right[c], c
        done.add(c)
        while cur in nxt:
            f = nxt[cur]
            seq += f[1][sk:]
            r = -10000000
            if prv[f[3]] is not f:
                break
            cur = f[3]
            if cur == c:
                break
            done.add(cur)
            seq += contigs[cur][sk:]
            r = right[cur]
        if not (left[c] <= -10000000 and r <= -10000000) and len(seq) >= min_contig:
            out.append((seq, left[c], r))

    for c in range(len(contigs)):
        if c not in prv:
            walk(c)
    for c in sorted(range(len(contigs)), key=lambda c: contigs[c][:k]):  # rings: opened at the smallest first k-mer
        if c not in done:
            walk(c)
    return sorted(out)


def reads_of(txt):
    a = np.frombuffer(txt, dtype=np.uint8)
    starts, lens = orc.fastq_reads(a, orc.FASTQ_RUN)
    return [txt[int(s):int(s) + int(n)].decode() for s, n in zip(starts, lens)]


# ---- the kernels on the host ------------------------------------------------------------------------------------------
class _EmuOut(C.Structure):
    _fields_ = [("n", C.c_uint64), ("off", C.POINTER(C.c_uint64)), ("bases", C.POINTER(C.c_char)), ("left", C.POINTER(C.c_int32)),
                ("right", C.POINTER(C.c_int32)), ("stats", C.c_uint64 * 6)]


def emu_stitch(contigs, left, right, txt, k, min_contig, chunk=0, cap0=4):
    path = os.path.join(ROOT, "tests", "hostemu", "libstitchemu.so")
    subprocess.run(["make", "-s", "-C", os.path.dirname(path)], check=True)
    E = C.CDLL(path)
    E.emu_stitch.restype = C.c_int
    E.emu_stitch.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                             C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(_EmuOut)]
    a = np.frombuffer(txt, dtype=np.uint8)
    starts, lens = orc.fastq_reads(a, orc.FASTQ_RUN)
    blob = np.frombuffer("".join(contigs).encode() + b"\0", dtype=np.uint8)
    offs = np.zeros(len(contigs) + 1, dtype=np.uint64)
    if contigs:
        offs[1:] = np.cumsum([len(c) for c in contigs])
    le, ri = np.ascontiguousarray(left, dtype=np.int32), np.ascontiguousarray(right, dtype=np.int32)
    o = _EmuOut()
    rc = E.emu_stitch(len(contigs), offs.ctypes.data, blob.ctypes.data, le.ctypes.data, ri.ctypes.data, a.ctypes.data, len(starts),
                      starts.ctypes.data, lens.ctypes.data, k, min_contig, chunk, cap0, C.byref(o))
    assert rc == 0, rc
    n = o.n
    of = np.ctypeslib.as_array(o.off, shape=(n + 1,)).copy()
    b = C.string_at(o.bases, int(of[-1])) if n else b""
    res = dict(contigs=[b[int(of[i]):int(of[i + 1])].decode() for i in range(n)], left=[o.left[i] for i in range(n)],
               right=[o.right[i] for i in range(n)], stats=list(o.stats))
    E.emu_stitch_free(C.byref(o))
    return res


# ---- hand-built cases ---------------------------------------------------------------------------------------------------
def hand_cases():
    rng = np.random.default_rng(77)

    def rnd(n):
        return "".join(rng.choice(list("ACGT"), n))

    k = 21
    A1, A2, B, Cc, D = rnd(120), rnd(130), rnd(140), rnd(100), rnd(60)
    cases = {}
    # two contigs bridged; a third whose flags are not "clean and covered <= 4 times"; a contig shorter than 61 bases
    cases["bridge"] = (k, [A1, B, Cc, D], [-2, -3, 7, -2], [-2, -4, -2, -2],
                       [A1[-45:] + "ACGTTGCAAC" + B[:50], orc.revcomp_str(A1[-30:] + "ACGTTGCAAC" + B[:44]), Cc[-40:] + "GG" + A1[:40], D[-30:] + "TT" + B[:30],
                        A1[-40:] + "ACGT" + Cc[:40]])
    # two fragments end on the same contig: the shorter joins it, the other record ends with its fragment
    cases["two arrive"] = (k, [A1, A2, B], [-2, -2, -2], [-2, -2, -2], [A1[-45:] + "ACGTTGCAAC" + B[:50], A2[-45:] + "TTG" + B[:50]])
    # several fragments leave one contig: the shortest stays, ties by sequence
    cases["two leave"] = (k, [A1, A2, B], [-2, -2, -2], [-2, -2, -2], [A1[-45:] + "ACGTTGCAAC" + B[:50], A1[-45:] + "ACGTAGCAAC" + B[:50], A1[-45:] + "ACGTTGCAACC" + A2[:50]])
    # the right probe of the contig the left probe came from is skipped (a read running from a contig's end into its own start)
    cases["own start"] = (k, [A1, B], [-2, -2], [-2, -2], [A1[-45:] + "ACGT" + A1[:40] + "CC" + B[:40], A1[-45:] + "ACGT" + A1[:40]])
    # ring: A1 -> B -> A1
    cases["ring"] = (k, [A1, B], [-2, -2], [-2, -2], [A1[-45:] + "ACGT" + B[:40], B[-45:] + "GGA" + A1[:40]])
    # probe collision: A2x ends with the (k-1)-mer A1 ends with
    A2x = A2[:-20] + A1[-20:]
    cases["collision"] = (k, [A1, A2x, B], [-2, -2, -2], [-2, -2, -2], [A1[-45:] + "ACGTTGCAAC" + B[:50]])
    # characters outside ACGT: forward they read as T, on the reverse strand complementary() gives 'N' -> T as well
    cases["n and lower case"] = (k, [A1, B], [-2, -2], [-2, -2], [A1[-45:] + "ACNTTgCAAC" + B[:50], orc.revcomp_str(A1[-45:] + "ACGTTGCAAC" + B[:50]).replace("G", "g", 1),
                                                                  orc.revcomp_str(A1[-45:]) .join(["", ""]) ])
    # reads too short to be looked at (readLength - (k-1) <= 1) and exactly long enough
    cases["short reads"] = (k, [A1, B], [-2, -2], [-2, -2], [A1[-20:] + B[:1], A1[-20:] + B[:2], (A1[-20:] + B[:20])])
    # several probes in one read: the first left-extendable hit and the LAST right-extendable hit of another contig count
    cases["many probes"] = (k, [A1, A2, B, Cc], [-2, -2, -2, -2], [-2, -2, -2, -2],
                            [A1[-30:] + "AC" + B[:25] + "GT" + A2[-25:] + "T" + Cc[:30], B[:25] + "GT" + A1[-30:] + "AC" + A1[:30] + "G" + Cc[:30] + A1[:25],
                             A1[-30:] + "AC" + A1[:30] + "G" + A1[:25]])
    # every case again with its reads handed over as the other strand: the same fragments must be found back to front
    for name, (kk, ctg, le, ri, reads) in list(cases.items()):
        if all(set(r) <= set("ACGT") for r in reads):
            cases[name + " (reads as the other strand)"] = (kk, ctg, le, ri, [orc.revcomp_str(r) for r in reads])
    return cases


def _as_fastq_case(case):
    k, contigs, left, right, reads = case
    return k, contigs, left, right, fq(reads)


# ---- CPU tests ----------------------------------------------------------------------------------------------------------
def test_stitch_oracle_bridges_low_coverage_gaps():
    G, txt = gap_scenario(5, 6000, 100, [(2000, 2060), (4000, 4070)])
    contigs, left, right = assembled(txt, 31)
    assert len(contigs) == 6 and all(-5 <= f < 0 for f in left + right)  # three pieces per strand, every end clean and thin
    res = orc.stitch(contigs, left, right, txt, 31, min_contig=500)
    assert res["stats"]["probes"] == 12 and res["stats"]["stitched_records"] == 2 and res["stats"]["rings"] == 0
    assert len(res["contigs"]) == 2
    for c in res["contigs"]:
        assert len(c) > 5900 and (c in G or orc.revcomp_str(c) in G)
    # -mincontig is applied to the stitched set: pieces shorter than it still take part
    res2 = orc.stitch(contigs, left, right, txt, 31, min_contig=5000)
    assert triples(res2) == triples(res)


def test_stitch_oracle_equals_the_literal_transcription_hand_cases():
    for name, case in hand_cases().items():
        k, contigs, left, right, txt = _as_fastq_case(case)
        got = orc.stitch(contigs, left, right, txt, k, min_contig=0)
        want = py_stitch(contigs, left, right, reads_of(txt), k, 0)
        assert triples(got) == want, name
    # what the cases are there for
    hc = hand_cases()
    k, contigs, left, right, txt = _as_fastq_case(hc["bridge"])
    res = orc.stitch(contigs, left, right, txt, k, 0)
    A1, B, Cc, D = contigs
    # Cc's right end is clean and thin, its left end is a fork winner (7): Cc -> A1 -> B is one record, D (60 bases) never probes
    assert triples(res) == sorted([(Cc + "GG" + A1 + "ACGTTGCAAC" + B, 7, -4), (D, -2, -2)])
    assert res["stats"]["probes"] == 5 and res["stats"]["stitched_records"] == 1
    k, contigs, left, right, txt = _as_fastq_case(hc["two arrive"])
    res = orc.stitch(contigs, left, right, txt, k, 0)
    assert (contigs[1] + "TTG" + contigs[2], -2, -2) in triples(res) and (contigs[0] + "ACGTTGCAAC" + contigs[2][:20], -2, -10000000) in triples(res)
    k, contigs, left, right, txt = _as_fastq_case(hc["ring"])
    res = orc.stitch(contigs, left, right, txt, k, 0)
    assert res["stats"]["rings"] == 1 and len(res["contigs"]) == 1 and res["right"].tolist() == [-10000000]
    k, contigs, left, right, txt = _as_fastq_case(hc["short reads"])
    assert orc.stitch(contigs, left, right, txt, k, 0)["stats"]["fragments"] == 1


def test_stitch_oracle_equals_the_literal_transcription_random():
    for k, txt in random_scenarios(12):
        contigs, left, right = assembled(txt, k)
        got = orc.stitch(contigs, left, right, txt, k, min_contig=0)
        assert triples(got) == py_stitch(contigs, left, right, reads_of(txt), k, 0)


def test_stitch_kernels_on_the_host_match_the_oracle():
    cases = [(k, *assembled(txt, k), txt) for k, txt in random_scenarios(16)]
    cases += [_as_fastq_case(c) for c in hand_cases().values()]
    cases.append((31, *assembled(gap_scenario(5, 6000, 100, [(2000, 2060), (4000, 4070)])[1], 31), gap_scenario(5, 6000, 100, [(2000, 2060), (4000, 4070)])[1]))
    rings = 0
    for k, contigs, left, right, txt in cases:
        for min_contig in (0, 200):
            want = orc.stitch(contigs, left, right, txt, k, min_contig=min_contig)
            for chunk, cap0 in ((0, 4), (7, 1), (100, 1000)):  # chunked scanning, the fragment list overflowing and not
                got = emu_stitch(contigs, left, right, txt, k, min_contig, chunk, cap0)
                assert triples(got) == triples(want)
                assert got["stats"] == list(want["stats"].values())
        rings += want["stats"]["rings"]
    assert rings > 0


def test_stitch_on_thinly_covered_reads_oracle_and_host_kernels():
    txt = natural_scenario(300_000, 8)
    contigs, left, right = assembled(txt, 31)
    want = orc.stitch(contigs, left, right, txt, 31, min_contig=0)
    st = want["stats"]
    assert len(contigs) > 500 and st["stitched_records"] > 100 and len(want["contigs"]) < 0.6 * len(contigs)
    assert max(len(c) for c in want["contigs"]) > 2 * max(len(c) for c in contigs)
    got = emu_stitch(contigs, left, right, txt, 31, 0, chunk=5000, cap0=64)
    assert triples(got) == triples(want) and got["stats"] == list(st.values())
    from workload import synth
    G = synth.genome(300_000, seed=7).tobytes().decode()
    for c in want["contigs"]:  # error-free reads: whatever was stitched is still a piece of the genome
        assert c in G or orc.revcomp_str(c) in G


def test_stitch_dense_probes_three_way():
    """Tiny k: nearly every window of a read is some contig end, so the first-left / last-right-of-another-contig rule, probe
    collisions and reads full of N are exercised on every read; oracle == literal transcription == kernels on the host."""
    frags = 0
    for seed in range(40):
        rng = np.random.default_rng(seed)
        k = int(rng.choice([4, 5, 6, 7]))
        nc = int(rng.integers(3, 14))
        contigs = ["".join(rng.choice(list("ACGT"), int(rng.integers(61, 90)))) for _ in range(nc)]
        left = [int(rng.choice([-1, -2, -5, -6, 3])) for _ in range(nc)]
        right = [int(rng.choice([-1, -3, -5, -7, 2])) for _ in range(nc)]
        reads = ["".join(rng.choice(list("ACGTN" if seed % 4 == 0 else "ACGT"), int(rng.integers(k - 1, 60)))) for _ in range(40)]
        txt = fq(reads)
        want = orc.stitch(contigs, left, right, txt, k, 0)
        got = emu_stitch(contigs, left, right, txt, k, 0, chunk=9, cap0=2)
        assert triples(want) == triples(got) == py_stitch(contigs, left, right, reads_of(txt), k, 0), seed
        assert got["stats"] == list(want["stats"].values())
        frags += want["stats"]["fragments"]
    assert frags > 300
    # contigs that start with the same k-mer (no assembly has them, a random set at k = 4 does): the probe of the later contig
    # stays and a ring is opened at the earlier one, in all three
    for seed in (2001, 2139, 2955, 3741):
        rng = np.random.default_rng(seed)
        k = int(rng.choice([4, 5, 6, 7, 9]))
        nc = int(rng.integers(2, 16))
        contigs = ["".join(rng.choice(list("ACGT"), int(rng.integers(55, 95)))) for _ in range(nc)]
        left = [int(rng.choice([-1, -2, -5, -6, 3, -10000000])) for _ in range(nc)]
        right = [int(rng.choice([-1, -3, -5, -7, 2, -10000000])) for _ in range(nc)]
        reads = ["".join(rng.choice(list("ACGTNacgtu" if seed % 5 == 0 else "ACGT"), int(rng.integers(max(1, k - 2), 70)))) for _ in range(int(rng.integers(5, 60)))]
        txt = fq(reads)
        want = orc.stitch(contigs, left, right, txt, k, 0)
        got = emu_stitch(contigs, left, right, txt, k, 0)
        assert triples(want) == triples(got) == py_stitch(contigs, left, right, reads_of(txt), k, 0), seed
        assert got["stats"] == list(want["stats"].values())


def test_stitch_for_k_above_31_is_the_plain_assembly():
    """ReflexivDSMain64's stitch branch asks a Hashtable<List<Long>, Integer> for a Long (DSMain64:1562-1600 against :119-131): no
    read is ever cut.  The restatement keeps the contigs (the -mincontig rule still applies) and refuses k > 63."""
    k = 33
    _, contigs, left, right, txt = _as_fastq_case(hand_cases()["bridge"])
    res = orc.stitch(contigs, left, right, txt, k, min_contig=0)
    assert triples(res) == sorted(zip(contigs, left, right)) and res["stats"]["probes"] == 0 and res["stats"]["fragments"] == 0
    assert [len(c) for c in orc.stitch(contigs, left, right, txt, k, min_contig=100)["contigs"]] == [len(c) for c in contigs if len(c) >= 100]
    with pytest.raises(ValueError):
        orc.stitch(["A" * 100], [-2], [-2], fq(["A" * 80]), 64)


# ---- GPU tests (through the C ABI) ----------------------------------------------------------------------------------------
def _gpu_stitch(txt, k, cover=3, min_contig=0, via_kmerc=False, **param):
    import reflexiv_b200 as R
    p = R.DefaultParam(kmerSize=k, minKmerCoverage=cover, minContig=min_contig, **param)
    with R.ReflexivContext(p, device=0) as ctx:
        ctx.push_fastq(txt)
        ctx.count()
        if via_kmerc:  # the -kmerc seam: the table comes back from the host
            keys, cnt = ctx.counts()
            ctx.reset()
            ctx.load_counts(keys, cnt)
        st0 = ctx.stitch_begin()
        ctx.push_fastq(txt)
        st = ctx.stitch_finish()
        got = ctx.contigs()
    assert st0["n_probes"] == st["n_probes"]
    return sorted(got), st


@pytest.mark.gpu
def test_stitch_matches_oracle_on_gap_scenarios(monkeypatch):
    scen = [(31, gap_scenario(5, 6000, 100, [(2000, 2060), (4000, 4070)])[1]), (31, gap_scenario(6, 6000, 100, [(2000, 2060), (4000, 4070)], err=0.02)[1]),
            (21, gap_scenario(7, 3000, 80, [(1000, 1050), (2000, 2040)], circular=True)[1])] + list(random_scenarios(10))
    stitched = rings = 0
    for i, (k, txt) in enumerate(scen):
        contigs, left, right = assembled(txt, k)
        for min_contig in (0, 300):
            want = orc.stitch(contigs, left, right, txt, k, min_contig=min_contig)
            if i % 3 == 1:
                monkeypatch.setenv("RFX_FASTQ_CHUNK_BYTES", "4096")  # many chunks: fragments accumulate across them
            got, st = _gpu_stitch(txt, k, min_contig=min_contig, via_kmerc=bool(i % 2))
            monkeypatch.delenv("RFX_FASTQ_CHUNK_BYTES", raising=False)
            assert got == triples(want), (i, k, min_contig)
            assert [st["n_probes"], st["n_fragments"], st["n_after_pass1"], st["n_joined"], st["n_stitched"], st["n_rings"]] == list(want["stats"].values())
        stitched += want["stats"]["stitched_records"]
        rings += want["stats"]["rings"]
    assert stitched > 10 and rings > 0


@pytest.mark.gpu
def test_stitch_on_thinly_covered_reads(monkeypatch):
    for glen, cov, err in ((1_000_000, 8, 0.0), (400_000, 10, 0.005)):
        txt = natural_scenario(glen, cov, err=err)
        contigs, left, right = assembled(txt, 31)
        want = orc.stitch(contigs, left, right, txt, 31, min_contig=200)
        monkeypatch.setenv("RFX_FASTQ_CHUNK_BYTES", str(8 << 20))
        got, st = _gpu_stitch(txt, 31, min_contig=200)
        monkeypatch.delenv("RFX_FASTQ_CHUNK_BYTES", raising=False)
        assert got == triples(want)
        assert [st["n_probes"], st["n_fragments"], st["n_after_pass1"], st["n_joined"], st["n_stitched"], st["n_rings"]] == list(want["stats"].values())
        assert st["n_stitched"] > 100


@pytest.mark.gpu
def test_stitch_without_anything_to_stitch_returns_the_assembly():
    import gzip
    import reflexiv_b200 as R
    gold = os.path.join(ROOT, "tests", "golden")
    txt = b"".join(gzip.open(os.path.join(gold, f)).read() for f in ("paired_dat1.fq.gz", "paired_dat2.fq.gz"))
    got, st = _gpu_stitch(txt, 31, cover=3, min_contig=500)
    p = R.DefaultParam(kmerSize=31, minKmerCoverage=3)
    with R.ReflexivContext(p, device=0) as ctx:
        ctx.push_fastq(txt)
        ctx.count()
        ctx.assemble()
        plain = sorted(ctx.contigs())
    assert got == plain and st["n_stitched"] == 0 and st["n_reads"] > 0


@pytest.mark.gpu
def test_stitch_error_paths():
    import reflexiv_b200 as R
    from reflexiv_b200 import _lib
    txt = gap_scenario(5, 3000, 100, [(1500, 1560)])[1]
    with R.ReflexivContext(R.DefaultParam(kmerSize=31, minKmerCoverage=3), device=0) as ctx:
        with pytest.raises(R.RfxError) as e:
            ctx.stitch_finish()
        assert e.value.code == _lib.RFX_E_STATE
        with pytest.raises(R.RfxError) as e:
            ctx.stitch_begin()  # no table yet
        assert e.value.code == _lib.RFX_E_STATE
    with R.ReflexivContext(R.DefaultParam(kmerSize=31, minKmerCoverage=3), device=0) as ctx:
        ctx.push_fastq(txt)
        ctx.count()
        ctx.stitch_begin()
        ctx.assemble()  # abandons the open stage: its fragments would refer to the contigs it was opened on
        with pytest.raises(R.RfxError) as e:
            ctx.stitch_finish()
        assert e.value.code == _lib.RFX_E_STATE
        assert len(ctx.contigs()) > 0
    # k > 31: the reference's branch can never cut a read (ReflexivDSMain64.java:1562-1600): the stage is the plain assembly
    with R.ReflexivContext(R.DefaultParam(kmerSize=41, minKmerCoverage=3, minContig=100), device=0) as ctx:
        ctx.push_fastq(txt)
        ctx.count()
        ctx.assemble()
        plain = sorted(ctx.contigs())
        assert ctx.stitch_begin()["n_probes"] == 0
        ctx.push_fastq(txt)
        st = ctx.stitch_finish()
        assert sorted(ctx.contigs()) == plain and st["n_fragments"] == 0 and st["n_reads"] > 0 and len(plain) > 0


@pytest.mark.gpu
def test_stitch_through_both_drivers(tmp_path):
    import reflexiv_b200 as R
    k = 31
    G, txt = gap_scenario(5, 6000, 100, [(2000, 2060), (4000, 4070)])
    fqp = tmp_path / "reads.fq"
    fqp.write_bytes(txt)
    # the count table the reference's `counter` would have written
    p = R.DefaultParam(kmerSize=k, minKmerCoverage=3, inputFqPath=str(fqp), outputPath=str(tmp_path / "cnt"))
    R.Pipelines(p).reflexivDSCounterPipe()
    kmerc = str(tmp_path / "cnt" / f"Count_{k}" / "part-*.csv")
    contigs, left, right = assembled(txt, k)
    want = sorted(s for s, _, _ in triples(orc.stitch(contigs, left, right, txt, k, min_contig=500)))
    assert len(want) == 2

    def parse(path):
        seqs, cur = [], None
        for line in open(path).read().splitlines():
            if line.startswith(">"):
                cur = []
                seqs.append(cur)
            else:
                cur.append(line)
        return sorted("".join(s) for s in seqs)

    # Python mirror
    p2 = R.DefaultParam(kmerSize=k, minKmerCoverage=3, inputFqPath=str(fqp), inputKmerPath=kmerc, outputPath=str(tmp_path / "py"), stitch=True)
    st = R.Pipelines(p2).reflexivDSMainPipe()
    assert st["stitch"]["n_stitched"] == 2
    assert parse(tmp_path / "py" / f"Assemble_{k}" / "part-00000") == want
    # without -stitch the same command leaves the six pieces
    p3 = R.DefaultParam(kmerSize=k, minKmerCoverage=3, inputFqPath=str(fqp), inputKmerPath=kmerc, outputPath=str(tmp_path / "py0"))
    R.Pipelines(p3).reflexivDSMainPipe()
    assert len(parse(tmp_path / "py0" / f"Assemble_{k}" / "part-00000")) == 6
    # C++ driver
    exe = os.path.join(ROOT, "reflexiv_b200", "reflexiv")
    r = subprocess.run([exe, "run", "-fastq", str(fqp), "-kmerc", kmerc, "-kmer", str(k), "-cover", "3", "-stitch", "-outfile", str(tmp_path / "cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert parse(tmp_path / "cc" / f"Assemble_{k}" / "part-00000") == want
    assert "records stitched" in r.stdout + r.stderr
