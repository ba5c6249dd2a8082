"""Host-side checks of the arithmetic the CUDA kernels are built from (reflexiv_b200/csrc/rfx_core.h), driven by
tests/hostemu/hostemu.cpp and compared with the oracle.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from conftest import make_reads


def _pack(hostemu, txt, starts, lens, k, fc=0, ec=0):
    a = np.frombuffer(bytes(txt), dtype=np.uint8)
    n = len(starts)
    elen = np.zeros(n, np.uint32)
    woff = np.zeros(n, np.uint64)
    nw = hostemu.emu_pack_reads(a.ctypes.data, starts.ctypes.data, lens.ctypes.data, C.c_int64(n), k, fc, ec, elen.ctypes.data, woff.ctypes.data, None)
    words = np.zeros(nw + 8, np.uint64)
    hostemu.emu_pack_reads(a.ctypes.data, starts.ctypes.data, lens.ctypes.data, C.c_int64(n), k, fc, ec, elen.ctypes.data, woff.ctypes.data, words.ctypes.data)
    return elen, woff, words


def _records(hostemu, elen, woff, words, k, m, n_bins):
    n = hostemu.emu_partition(words.ctypes.data, C.c_int64(len(words)), elen.ctypes.data, woff.ctypes.data, C.c_int64(len(elen)), k, m, C.c_uint32(n_bins))
    recw = 2 if k <= 31 else 4
    recs = np.zeros((n, recw), np.uint64)
    bins = np.zeros(n, np.uint32)
    hostemu.emu_partition_fetch(recs.ctypes.data, bins.ctypes.data)
    return recs, bins


def _count_records(hostemu, recs, bins, k, m, n_bins):
    bad = C.c_int64(0)
    d = hostemu.emu_count_records(recs.ctypes.data, bins.ctypes.data, C.c_int64(len(recs)), k, m, C.c_uint32(n_bins), C.addressof(bad))
    hi = np.zeros(d, np.uint64); lo = np.zeros(d, np.uint64); cnt = np.zeros(d, np.uint32)
    hostemu.emu_count_fetch(hi.ctypes.data, lo.ctypes.data, cnt.ctypes.data)
    return hi, lo, cnt, bad.value


@pytest.mark.parametrize("k,m", [(31, 11), (31, 15), (21, 9), (15, 15), (61, 11), (33, 16), (5, 5)])
def test_records_reproduce_the_kmer_multiset(orc, hostemu, example_text, k, m):
    txt = example_text[:400_000]
    starts, lens = orc.fastq_reads(txt, orc.FASTQ_RUN)
    ref = orc.count_kmers(txt, starts, lens, k)
    elen, woff, words = _pack(hostemu, txt, starts, lens, k)
    n_bins = 96
    recs, bins = _records(hostemu, elen, woff, words, k, m, n_bins)
    hi, lo, cnt, bad = _count_records(hostemu, recs, bins, k, m, n_bins)
    assert bad == 0                                   # every record's first k-mer maps to the record's bin
    assert np.array_equal(hi, ref["keys_hi"]) and np.array_equal(lo, ref["keys_lo"]) and np.array_equal(cnt, ref["counts"])
    nk = recs[:, 0] >> np.uint64(48)
    recw = recs.shape[1]
    assert nk.min() >= 1 and nk.max() <= (recw * 64 - 16) // 2 - k + 1
    assert int(nk.sum()) == ref["n_instances"]
    assert bins.max() < n_bins


@pytest.mark.parametrize("k,m", [(31, 11), (25, 7), (61, 13)])
def test_sliding_minimum_equals_brute_force(orc, hostemu, k, m):
    """van Herk on-the-fly window minimum vs recomputing the minimiser of every k-mer from scratch, and the
    strand symmetry that makes the bin a function of the canonical k-mer."""
    rng = np.random.default_rng(k * 100 + m)
    n_bins = 1000
    for length in (k, k + 1, k + 7, 100, 151, 300):
        seq = "".join("ACGT"[i] for i in rng.integers(0, 4, length))
        for s in (seq, orc.revcomp_str(seq)):
            txt = s.encode() + b"\n"
            st = np.array([0], np.uint64); ln = np.array([length], np.uint32)
            # pack with the k > 31 rule so reads of length k are kept for every k here
            elen, woff, words = _pack(hostemu, txt, st, ln, max(k, 33) if length - k <= 1 else k)
            elen[0] = length
            brute = np.zeros(length - k + 1, np.uint32)
            hostemu.emu_bins_bruteforce(words.ctypes.data, C.c_uint32(length), k, m, C.c_uint32(n_bins), brute.ctypes.data)
            recs, bins = _records(hostemu, elen, woff, words, k, m, n_bins)
            nk = (recs[:, 0] >> np.uint64(48)).astype(np.int64)
            expanded = np.repeat(bins, nk)
            assert np.array_equal(expanded, brute)
            if s is seq:
                fwd_bins = brute
            else:
                assert np.array_equal(brute, fwd_bins[::-1])   # k-mer i of the reverse strand is k-mer n-1-i


def test_packing_matches_clip_and_length_rules(orc, hostemu):
    txt = b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTACGTAC\nACGTACG\n" + b"N" * 40 + b"\n"
    starts, lens = orc.fastq_reads(txt, orc.FASTQ_LINE)
    for k, fc, ec in ((7, 0, 0), (7, 3, 2), (5, 1, 0), (33, 0, 0), (33, 5, 5)):
        elen, woff, words = _pack(hostemu, txt, starts, lens, k, fc, ec)
        ref = orc.count_kmers(txt, starts, lens, k, fc, ec)
        inst = int(sum(max(0, int(e) - k + 1) for e in elen))
        assert inst == ref["n_instances"]
        recs, bins = _records(hostemu, elen, woff, words, k, min(k, 11), 64)
        hi, lo, cnt, bad = _count_records(hostemu, recs, bins, k, min(k, 11), 64)
        assert bad == 0 and np.array_equal(lo, ref["keys_lo"]) and np.array_equal(hi, ref["keys_hi"]) and np.array_equal(cnt, ref["counts"])


def test_revcomp_bit_tricks(orc, hostemu):
    rng = np.random.default_rng(1)
    for nb in (1, 2, 5, 15, 16, 30, 31, 32):
        for _ in range(20):
            s = "".join("ACGT"[i] for i in rng.integers(0, 4, nb))
            v = int("".join(format("ACGT".index(c), "02b") for c in s), 2)
            r = int("".join(format("ACGT".index(c), "02b") for c in orc.revcomp_str(s)), 2)
            assert hostemu.emu_revcomp64(v, nb) == r
    for nb in (33, 40, 61, 63, 64):
        s = "".join("ACGT"[i] for i in rng.integers(0, 4, nb))
        v = int("".join(format("ACGT".index(c), "02b") for c in s), 2)
        r = int("".join(format("ACGT".index(c), "02b") for c in orc.revcomp_str(s)), 2)
        oh, ol = C.c_uint64(), C.c_uint64()
        hostemu.emu_revcomp128(C.c_uint64(v >> 64), C.c_uint64(v & (2**64 - 1)), nb, C.addressof(oh), C.addressof(ol))
        assert (oh.value << 64) | ol.value == r


def _emu_fork(hostemu, cnt, k, E):
    n = hostemu.emu_fork_filter(cnt["keys_hi"].ctypes.data, cnt["keys_lo"].ctypes.data, cnt["counts"].ctypes.data, C.c_int64(len(cnt["counts"])), k, E)
    hi = np.zeros(n, np.uint64); lo = np.zeros(n, np.uint64); le = np.zeros(n, np.int32); ri = np.zeros(n, np.int32)
    hostemu.emu_fork_fetch(hi.ctypes.data, lo.ctypes.data, le.ctypes.data, ri.ctypes.data)
    order = np.lexsort((lo, hi))
    return hi[order], lo[order], le[order], ri[order]


@pytest.mark.parametrize("k,E,cover,err", [(31, 8, 2, 0.01), (31, 0, 2, 0.01), (21, 8, 1, 0.02), (15, 8, 1, 0.03), (41, 8, 2, 0.01), (12, 8, 1, 0.02)])
def test_fork_filters_per_group_formulation(orc, hostemu, k, E, cover, err):
    """The GPU evaluates each (k-1)-mer group from its four candidate neighbours (right_fork / left_fork in rfx_core.h);
    the oracle sorts and scans like the reference.  Both must keep the same oriented k-mers with the same flags,
    including real forks (repeats), ties and, for even k, palindromes."""
    from workload import synth
    g = synth.genome(3000, 7 + k)
    g[1500:1900] = g[200:600]          # a repeat -> real forks
    g[2500:2520] = np.frombuffer(b"ACGT" * 5, np.uint8)  # low-complexity / palindromic stretch
    txt = synth.fastq(g, 900, read_len=100, frag_len=250, error_rate=err, seed_reads=3, seed_errors=4)
    starts, lens = orc.fastq_reads(txt, orc.FASTQ_RUN)
    cnt = orc.count_kmers(txt, starts, lens, k, min_count=cover)
    ref = orc.fork_filter(cnt["keys_hi"], cnt["keys_lo"], cnt["counts"], k, E)
    hi, lo, le, ri = _emu_fork(hostemu, cnt, k, E)
    assert np.array_equal(hi, ref["keys_hi"]) and np.array_equal(lo, ref["keys_lo"])
    assert np.array_equal(le, ref["left"]) and np.array_equal(ri, ref["right"])
    if k == 31 and E == 8:
        assert ref["stats"]["right_forks"] > 0


def _emu_sorted(hostemu, cnt, k, E, fold, X):
    n = hostemu.emu_sorted_filter(cnt["keys_hi"].ctypes.data, cnt["keys_lo"].ctypes.data, cnt["counts"].ctypes.data, C.c_int64(len(cnt["counts"])), k, E,
                                  C.c_double(fold), X)
    hi = np.zeros(n, np.uint64); lo = np.zeros(n, np.uint64); le = np.zeros(n, np.int32); ri = np.zeros(n, np.int32)
    hostemu.emu_fork_fetch(hi.ctypes.data, lo.ctypes.data, le.ctypes.data, ri.ctypes.data)
    order = np.lexsort((lo, hi))
    return hi[order], lo[order], le[order], ri[order]


@pytest.mark.parametrize("k,E,fold,kmax,cover,err", [(31, 8, 1.5, 95, 1, 0.01), (31, 6, 2.0, 31, 2, 0.01), (23, 8, 1.5, 95, 1, 0.02), (15, 3, 1.5, 53, 1, 0.03),
                                                      (41, 8, 1.5, 95, 1, 0.01), (12, 8, 1.5, 95, 1, 0.02), (53, 24, 2.0, 95, 1, 0.01)])
def test_sorted_stage_per_group_formulation(orc, hostemu, k, E, fold, kmax, cover, err):
    """Count_<k>_sorted: sorted_right_fork / sorted_left_fork evaluated per (k-1)-mer group, as the GPU does,
    against the oracle's sort + sequential scan with the reference's packed attribute word -- including the two places
    where the reference lets a weaker row's coverage / right flag leak into the stored row."""
    from workload import synth
    g = synth.genome(3000, 17 + k)
    g[1500:1900] = g[200:600]
    g[2500:2520] = np.frombuffer(b"ACGT" * 5, np.uint8)
    txt = synth.fastq(g, 900, read_len=100, frag_len=250, error_rate=err, seed_reads=5, seed_errors=6)
    starts, lens = orc.fastq_reads(txt, orc.FASTQ_RUN)
    cnt = orc.count_kmers(txt, starts, lens, k, min_count=cover)
    cnt["counts"][::97] = 40000 + np.arange(len(cnt["counts"][::97]), dtype=np.uint32)   # above the 30000 saturation
    ref = orc.sorted_rows(cnt["keys_hi"], cnt["keys_lo"], cnt["counts"], k, E, fold, kmax)
    hi, lo, le, ri = _emu_sorted(hostemu, cnt, k, E, fold, kmax + 3)
    assert np.array_equal(hi, ref["keys_hi"]) and np.array_equal(lo, ref["keys_lo"])
    assert np.array_equal(le, ref["left"]) and np.array_equal(ri, ref["right"])
    assert set(np.unique(le)) <= {-1, kmax + 3} and set(np.unique(ri)) <= {-1, kmax + 3}
    if k == 31:
        assert (le == kmax + 3).any() and (ri == kmax + 3).any()     # real forks are present


def test_sorted_stage_reference_quirks_by_hand(orc):
    """Three hand-made groups (k = 5, E = 2, fold 1.5, list maximum 5 -> X = 8):
       * prefix AAAA with last bases A:5, C:3 -- neither is an error: A survives but carries C's coverage 3 on
         (LeftAndRightSorting.java:504-518), which then decides the left filter of its suffix group;
       * coverage 1 never wins a fork flag (highestLeftMarker == 1 -> -1)."""
    def enc(s):
        v = 0
        for ch in s:
            v = (v << 2) | "ACGT".index(ch)
        return v
    def canon(s):
        rc = s.translate(str.maketrans("ACGT", "TGCA"))[::-1]
        return min(s, rc)
    rows = {"AAAAA": 5, "AAAAC": 3, "CAAAA": 4, "GGGCA": 1, "GGGCC": 1}
    table = {}
    for s, c in rows.items():
        table[canon(s)] = c
    keys = np.array([enc(s) for s in table], np.uint64)
    res = orc.sorted_rows(np.zeros(len(keys), np.uint64), keys, np.array(list(table.values()), np.uint32), 5, 2, 1.5, 5)
    got = {orc.decode_kmer(0, l, 5): (int(a), int(b)) for l, a, b in zip(res["keys_lo"], res["left"], res["right"])}
    # right filter, prefix AAAA: A(5) then C(3): C is no error (3 > E) -> A stays with coverage 3, right = 8.
    # left filter, suffix AAAA: first bases A (AAAAA, carried coverage 3) then C (CAAAA, coverage 4): 4 > 3, 3 > E -> CAAAA wins
    # with left = 8 although AAAAA was counted 5 times.
    assert "AAAAA" not in got and got["CAAAA"][0] == 8
    # GGGC + A / C, both coverage 1: the larger base wins, flag -1 because the coverage is 1
    assert "GGGCA" not in got and got["GGGCC"][1] == -1


def test_canonical_assembly_vs_pass_simulation_on_clean_data(orc):
    """Error-free reads, no repeats: no budget flag survives, so the fixed point is order independent.  The
    reference's pass simulation may stop before it (its stopping rule only compares record counts three passes
    apart, DSMain.java:297-311), so its contigs must be substrings of the fixed-point contigs and together cover
    them."""
    txt = make_reads(11, 6000, 2400, read_len=100, err=0.0)
    a = orc.run_pipeline(txt, k=31, cover=2, min_contig=31, mode=orc.ASM_CANONICAL)
    b = orc.run_pipeline(txt, k=31, cover=2, min_contig=31, mode=orc.ASM_REFSIM)
    assert a["asm"]["n_budget_junctions"] == 0
    fix = a["asm"]["contigs"]
    assert max(len(c) for c in fix) > 5000
    for c in b["asm"]["contigs"]:
        assert any(c in f for f in fix)
    # same k-mer content: (#bases - (k-1)) summed over contigs = number of oriented k-mers
    assert sum(len(c) - 30 for c in fix) == sum(len(c) - 30 for c in b["asm"]["contigs"]) == len(a["forks"]["left"])


@pytest.mark.parametrize("k,m", [(31, 11), (21, 9), (5, 5), (32, 11), (33, 16), (47, 11), (61, 11), (63, 13)])
def test_flat_kmer_cut_equals_rolling_extraction(orc, hostemu, k, m):
    """rec_kmer_at -- what a lane of the counting kernel's warp-wide expansion does -- against rec_foreach_kmer, on records
    cut from long reads so that k = 32 reaches the last word of a 32-byte record (89 k-mers: bit offset 192)."""
    txt = bytes(make_reads(5 + k, 3_000, 60, read_len=400, err=0.0, frag=600))
    starts, lens = orc.fastq_reads(txt, orc.FASTQ_RUN)
    elen, woff, words = _pack(hostemu, txt, starts, lens, k)
    recs, _ = _records(hostemu, elen, woff, words, k, m, 1)     # one bin: a read is one run, cut into full records
    nk = (recs[:, 0] >> np.uint64(48)).astype(np.int64)
    recw = recs.shape[1]
    assert nk.max() == (recw * 64 - 16) // 2 - k + 1            # records of maximal length are among them
    assert hostemu.emu_check_kmer_at(recs.ctypes.data, C.c_int64(len(recs)), k) == 0


def test_table_hashes_do_not_cancel_top_bits(hostemu):
    """Two 61-mers from the reference's example that differ in the top bits of two 32-bit words collided in BOTH hashes
    of a plain (word * odd) ^ ... combination; the folded multiplies separate them, and close relatives in general."""
    out = (C.c_uint32 * 4)()
    def h(hi, lo):
        hostemu.emu_table_hashes(C.c_uint64(hi), C.c_uint64(lo), out)
        return tuple(out)
    a = h(0x003bfa4dbcc08ba3, 0xdb2e38efd1733f32)
    b = h(0x003bfa4dfcc08ba3, 0xdb2e38ef11733f32)
    assert a[0] != b[0] and a[1] != b[1]
    rng = np.random.default_rng(11)
    seen_pair, seen_narrow = set(), set()
    n = 0
    for _ in range(300):
        hi, lo = int(rng.integers(0, 1 << 58)), int(rng.integers(0, 1 << 63))
        for w in range(4):                       # flip the top two bits of every 32-bit word, alone and in pairs
            for w2 in range(w, 4):
                dhi = ((3 << 30) << (32 * (w - 2))) if w >= 2 else 0
                dlo = ((3 << 30) << (32 * w)) if w < 2 else 0
                dhi ^= ((1 << 30) << (32 * (w2 - 2))) if w2 >= 2 else 0
                dlo ^= ((1 << 30) << (32 * w2)) if w2 < 2 else 0
                r = h(hi ^ dhi, lo ^ dlo)
                seen_pair.add((r[0], r[1])); seen_narrow.add((r[2], r[3]))
                n += 1
    assert len(seen_pair) == n                   # no pair of relatives shares both table hashes
    # the narrow hash only sees lo: of the ten variants per key, (0,2)/(0,3), (1,2)/(1,3) and (2,2)/(2,3)/(3,3) share lo
    assert len(seen_narrow) == 300 * 6


def _random_table(rng, k, density, max_count, heavy=0.0):
    """A random canonical count table over ALL 4^k k-mers (small k): dense enough that 3- and 4-way forks, coverage ties
    and (even k) palindromes are everywhere."""
    n_all = 1 << (2 * k)
    v = np.arange(n_all, dtype=np.uint64)
    rc = np.zeros(n_all, dtype=np.uint64)
    x = v.copy()
    for _ in range(k):
        rc = (rc << np.uint64(2)) | (np.uint64(3) - (x & np.uint64(3)))
        x >>= np.uint64(2)
    canon = v[v <= rc]
    keep = canon[rng.random(len(canon)) < density]
    counts = rng.integers(1, max_count + 1, len(keep)).astype(np.uint32)
    if heavy:
        big = rng.random(len(keep)) < heavy
        counts[big] = rng.integers(29_990, 30_020, int(big.sum())).astype(np.uint32)   # around the 30000 saturation of the sorted stage
    return dict(keys_hi=np.zeros(len(keep), np.uint64), keys_lo=keep, counts=counts)


@pytest.mark.parametrize("k", [3, 4, 5, 6, 7])
def test_fork_filters_fuzz_on_dense_random_tables(orc, hostemu, k):
    """Both fork-filter pairs (the `run` command's and the sorted stage's), per-group GPU formulation vs the oracle's
    sort + scan, on tables where most (k-1)-mer groups hold 2-4 candidates."""
    rng = np.random.default_rng(100 + k)
    for trial in range(12):
        cnt = _random_table(rng, k, density=float(rng.choice([0.15, 0.5, 0.9])), max_count=int(rng.choice([2, 5, 20])), heavy=0.05 if trial % 3 == 0 else 0.0)
        if len(cnt["counts"]) == 0:
            continue
        for E in (0, 3, 8):
            ref = orc.fork_filter(cnt["keys_hi"], cnt["keys_lo"], cnt["counts"], k, E)
            hi, lo, le, ri = _emu_fork(hostemu, cnt, k, E)
            assert np.array_equal(lo, ref["keys_lo"]) and np.array_equal(le, ref["left"]) and np.array_equal(ri, ref["right"]), (k, trial, E)
        for E, fold, kmax in ((2, 1.5, 7), (8, 2.0, 95), (1, 1.0, 31)):
            if (k - 1) % 31 == 0:
                continue
            ref = orc.sorted_rows(cnt["keys_hi"], cnt["keys_lo"], cnt["counts"], k, E, fold, kmax)
            hi, lo, le, ri = _emu_sorted(hostemu, cnt, k, E, fold, kmax + 3)
            assert np.array_equal(lo, ref["keys_lo"]) and np.array_equal(le, ref["left"]) and np.array_equal(ri, ref["right"]), (k, trial, E, fold)


def test_newline_masks_equal_a_bytewise_scan():
    """K1, rfx_newline.cuh: the per-chunk newline mask and the newline-followed-by-'@' mask (what pass 2 of the line scan works
    from instead of the text) against a byte-by-byte scan: random texts, every alignment, garbage in front and behind."""
    import ctypes as C
    import os
    import subprocess
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    E = C.CDLL(os.path.join(d, "libnewlineemu.so"))
    E.emu_newline_masks.restype = C.c_int64
    E.emu_newline_masks.argtypes = [C.c_uint32, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
    lines = C.c_int64()
    assert E.emu_newline_masks(3, 30000, 700, C.byref(lines)) == 0
    assert lines.value > 1_000_000
    assert E.emu_newline_masks(4, 2000, 5, C.byref(lines)) == 0  # texts shorter than a chunk, empty texts
