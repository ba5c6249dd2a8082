"""ctypes loader for the CPU oracle (oracle/reflexiv_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, from
``__graft_entry__.smoke()`` and from the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py``.  Nothing under ``reflexiv_b200/`` imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    """Compile liboracle.so in place (gcc, seconds)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return os.path.join(_HERE, "liboracle.so")


class _Contigs(C.Structure):
    _fields_ = [
        ("n_contigs", C.c_int64),
        ("offsets", C.POINTER(C.c_uint64)),
        ("bases", C.POINTER(C.c_char)),
        ("left", C.POINTER(C.c_int32)),
        ("right", C.POINTER(C.c_int32)),
        ("n_passes", C.c_int64),
        ("n_budget_junctions", C.c_int64),
        ("n_budget_admissible", C.c_int64),
        ("n_cycles", C.c_int64),
    ]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "reflexiv_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        L.orc_fastq_reads.restype = C.c_int64
        L.orc_fastq_reads.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.orc_count_kmers.restype = C.c_int64
        L.orc_count_kmers.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_fork_filter.restype = C.c_int64
        L.orc_fork_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p), C.c_void_p]
        L.orc_sorted_rows.restype = C.c_int64
        L.orc_sorted_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int,
                                      C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p)]
        L.orc_assemble.restype = C.c_int
        L.orc_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.POINTER(_Contigs)]
        L.orc_stitch.restype = C.c_int
        L.orc_stitch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_int, C.c_int, C.POINTER(_Contigs), C.c_void_p]
        L.orc_contigs_free.argtypes = [C.POINTER(_Contigs)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _take(ptr: C.c_void_p, n: int, dtype) -> np.ndarray:
    """Copy a malloc'ed C array into numpy and free it."""
    dt = np.dtype(dtype)
    if n == 0:
        out = np.empty(0, dtype=dt)
    else:
        buf = (C.c_char * (n * dt.itemsize)).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=dt, count=n).copy()
    lib().orc_free(ptr)
    return out


def _as_bytes_array(txt) -> np.ndarray:
    if isinstance(txt, np.ndarray):
        return np.ascontiguousarray(txt, dtype=np.uint8)
    return np.frombuffer(bytes(txt), dtype=np.uint8)


FASTQ_RUN, FASTQ_COUNTER, FASTQ_LINE = 0, 1, 2
ASM_CANONICAL, ASM_REFSIM, ASM_SCHEDULED = 0, 1, 2


def set_threads(n: int):
    """Threads for the sort-based stages (fork filters, REFSIM passes); 1 = the deterministic single-partition order."""
    lib().orc_set_threads(int(n))


def fastq_reads(txt, mode: int = FASTQ_RUN):
    """A1 / A1': returns (starts uint64[n], lens uint32[n]) of the sequence lines kept."""
    a = _as_bytes_array(txt)
    ps, pl = C.c_void_p(), C.c_void_p()
    n = lib().orc_fastq_reads(a.ctypes.data, a.size, mode, C.byref(ps), C.byref(pl))
    return _take(ps, n, np.uint64), _take(pl, n, np.uint32)


def count_kmers(txt, starts, lens, k: int, front_clip: int = 0, end_clip: int = 0,
                min_count: int = 1, max_count: int = 2**62, n_threads: int = 1):
    """A2-A4: returns dict(keys_hi, keys_lo, counts, n_instances, n_distinct); rows sorted by key."""
    a = _as_bytes_array(txt)
    starts = np.ascontiguousarray(starts, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    ph, pl, pc = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ni, nd = C.c_int64(), C.c_int64()
    n = lib().orc_count_kmers(a.ctypes.data, starts.ctypes.data, lens.ctypes.data, len(starts), k, front_clip,
                              end_clip, min_count, max_count, n_threads, C.byref(ph), C.byref(pl), C.byref(pc),
                              C.byref(ni), C.byref(nd))
    if n < 0:
        raise ValueError(f"orc_count_kmers failed: {n}")
    return dict(keys_hi=_take(ph, n, np.uint64), keys_lo=_take(pl, n, np.uint64), counts=_take(pc, n, np.uint32),
                n_instances=ni.value, n_distinct=nd.value)


FORK_STATS = ("right_forks", "right_forks_ge3", "right_ties", "left_forks", "left_forks_ge3", "left_ties",
              "left_flag_nonneg", "right_flag_nonneg")


def fork_filter(keys_hi, keys_lo, counts, k: int, min_error_cov: int = 8):
    """A6-A8: returns dict(keys_hi, keys_lo, left, right, stats) of surviving oriented k-mers, sorted by key."""
    kh = np.ascontiguousarray(keys_hi, dtype=np.uint64)
    kl = np.ascontiguousarray(keys_lo, dtype=np.uint64)
    ct = np.ascontiguousarray(counts, dtype=np.uint32)
    ph, pl, pL, pR = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    stats = np.zeros(8, dtype=np.int64)
    n = lib().orc_fork_filter(kh.ctypes.data, kl.ctypes.data, ct.ctypes.data, len(kh), k, min_error_cov,
                              C.byref(ph), C.byref(pl), C.byref(pL), C.byref(pR), stats.ctypes.data)
    if n < 0:
        raise ValueError(f"orc_fork_filter failed: {n}")
    return dict(keys_hi=_take(ph, n, np.uint64), keys_lo=_take(pl, n, np.uint64), left=_take(pL, n, np.int32),
                right=_take(pR, n, np.int32), stats=dict(zip(FORK_STATS, stats.tolist())))


def sorted_rows(keys_hi, keys_lo, counts, k: int, min_error_cov: int = 8, min_repeat_fold: float = 1.5,
                max_kmer_size: int = 95, max_cov: int = 10_000_000):
    """the rows of Count_<k>_sorted as dict(keys_hi, keys_lo, left, right), sorted by key."""
    kh = np.ascontiguousarray(keys_hi, dtype=np.uint64)
    kl = np.ascontiguousarray(keys_lo, dtype=np.uint64)
    ct = np.ascontiguousarray(counts, dtype=np.uint32)
    ph, pl, pL, pR = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = lib().orc_sorted_rows(kh.ctypes.data, kl.ctypes.data, ct.ctypes.data, len(kh), k, min_error_cov, min_repeat_fold,
                              max_kmer_size, max_cov, C.byref(ph), C.byref(pl), C.byref(pL), C.byref(pR))
    if n < 0:
        raise ValueError(f"orc_sorted_rows: outside the reference's working domain ({n})")
    return dict(keys_hi=_take(ph, n, np.uint64), keys_lo=_take(pl, n, np.uint64), left=_take(pL, n, np.int32),
                right=_take(pR, n, np.int32))


def sorted_rows_text(res, k: int) -> str:
    """`KMER,1|left|right\n` rows (DSBinaryFullKmerArrayToString, LeftAndRightSorting.java:249-274), sorted by key."""
    return "".join(f"{decode_kmer(h, l, k)},1|{a}|{b}\n" for h, l, a, b in zip(res["keys_hi"], res["keys_lo"], res["left"], res["right"]))


def assemble(keys_hi, keys_lo, left, right, k: int, min_contig: int = 500, mode: int = ASM_CANONICAL,
             min_iter: int = 15, max_iter: int = 150):
    """A9-A10: returns dict(contigs=[str], left, right, n_passes, n_budget_junctions, n_budget_admissible, n_cycles)."""
    kh = np.ascontiguousarray(keys_hi, dtype=np.uint64)
    kl = np.ascontiguousarray(keys_lo, dtype=np.uint64)
    le = np.ascontiguousarray(left, dtype=np.int32)
    ri = np.ascontiguousarray(right, dtype=np.int32)
    c = _Contigs()
    rc = lib().orc_assemble(kh.ctypes.data, kl.ctypes.data, le.ctypes.data, ri.ctypes.data, len(kh), k, min_contig,
                            mode, min_iter, max_iter, C.byref(c))
    if rc != 0:
        raise ValueError(f"orc_assemble failed: {rc}")
    n = c.n_contigs
    offs = np.ctypeslib.as_array(c.offsets, shape=(n + 1,)).copy()
    total = int(offs[-1])
    blob = C.string_at(c.bases, total)
    out = dict(contigs=[blob[int(offs[i]):int(offs[i + 1])].decode() for i in range(n)],
               left=np.ctypeslib.as_array(c.left, shape=(n,)).copy() if n else np.empty(0, np.int32),
               right=np.ctypeslib.as_array(c.right, shape=(n,)).copy() if n else np.empty(0, np.int32),
               n_passes=c.n_passes, n_budget_junctions=c.n_budget_junctions,
               n_budget_admissible=c.n_budget_admissible, n_cycles=c.n_cycles)
    lib().orc_contigs_free(C.byref(c))
    return out


def _contigs_out(c: _Contigs) -> dict:
    n = c.n_contigs
    offs = np.ctypeslib.as_array(c.offsets, shape=(n + 1,)).copy()
    blob = C.string_at(c.bases, int(offs[-1])) if n else b""
    return dict(contigs=[blob[int(offs[i]):int(offs[i + 1])].decode() for i in range(n)],
                left=np.ctypeslib.as_array(c.left, shape=(n,)).copy() if n else np.empty(0, np.int32),
                right=np.ctypeslib.as_array(c.right, shape=(n,)).copy() if n else np.empty(0, np.int32))


STITCH_STATS = ("probes", "fragments", "after_pass1", "joined_both_sides", "stitched_records", "rings")


def stitch(contigs, left, right, txt, k: int, min_contig: int = 500):
    """SURVEY 8f-4 (`-stitch`, ReflexivDSMain.java:585-672): `contigs` are ALL records of the extension (assemble with
    min_contig=0) with their flags; `txt` is the FASTQ text, read through the `run` filter.  Returns dict(contigs, left,
    right, stats)."""
    a = _as_bytes_array(txt)
    starts, lens = fastq_reads(a, FASTQ_RUN)
    blob = "".join(contigs).encode()
    offs = np.zeros(len(contigs) + 1, dtype=np.uint64)
    if contigs:
        offs[1:] = np.cumsum([len(c) for c in contigs])
    le = np.ascontiguousarray(left, dtype=np.int32)
    ri = np.ascontiguousarray(right, dtype=np.int32)
    bb = np.frombuffer(blob + b"\0", dtype=np.uint8)
    stats = np.zeros(6, dtype=np.int64)
    c = _Contigs()
    rc = lib().orc_stitch(len(contigs), offs.ctypes.data, bb.ctypes.data, le.ctypes.data, ri.ctypes.data, a.ctypes.data,
                          starts.ctypes.data, lens.ctypes.data, len(starts), k, min_contig, C.byref(c), stats.ctypes.data)
    if rc != 0:
        raise ValueError(f"orc_stitch: k = {k} is outside the reference's range ({rc})")
    out = _contigs_out(c)
    out["stats"] = dict(zip(STITCH_STATS, stats.tolist()))
    lib().orc_contigs_free(C.byref(c))
    return out


# ---- small pure-Python helpers shared by tests ------------------------------------------------

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp_str(s: str) -> str:
    return s.encode().translate(_COMP)[::-1].decode()


def canonical_contig_set(contigs):
    """Contigs in canonical orientation (min of strand / reverse complement), sorted: the comparison
    domain for A10 (every contig is emitted on both strands, order is partition dependent)."""
    return sorted(min(c, revcomp_str(c)) for c in contigs)


def decode_kmer(hi: int, lo: int, k: int) -> str:
    v = (int(hi) << 64) | int(lo)
    return "".join("ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def count_table_text(res, k: int) -> str:
    """`KMER,count\\n` rows, the reference's CSV row format (Counter:405-428), sorted by key."""
    return "".join(f"{decode_kmer(h, l, k)},{c}\n" for h, l, c in zip(res["keys_hi"], res["keys_lo"], res["counts"]))


def run_pipeline(txt, k: int = 31, cover: int = 2, maxcov: int = 10_000_000, min_error_cov: int = 8,
                 min_contig: int = 500, mode: int = ASM_CANONICAL, front_clip: int = 0, end_clip: int = 0,
                 fastq_mode: int = FASTQ_RUN, n_threads: int = 1):
    """`reflexiv run` end to end on FASTQ text (ReflexivDSMain.assembly, DSMain:123-357)."""
    starts, lens = fastq_reads(txt, fastq_mode)
    cnt = count_kmers(txt, starts, lens, k, front_clip, end_clip, cover, maxcov, n_threads)
    ff = fork_filter(cnt["keys_hi"], cnt["keys_lo"], cnt["counts"], k, min_error_cov)
    asm = assemble(ff["keys_hi"], ff["keys_lo"], ff["left"], ff["right"], k, min_contig, mode)
    return dict(n_reads=len(starts), counts=cnt, forks=ff, asm=asm)
