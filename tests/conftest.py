import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

# Ranks that share a device (tests/test_multi_gpu.py) must not share a hardware work queue: a rank waiting in a cross-GPU
# barrier would hold up the kernels of the rank it is waiting for.  Read by the driver when CUDA starts.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and no kernel may be loaded lazily while a peer rank on the SAME device spins in a barrier (loading a module can wait
# for the device to drain): load every kernel when the library is loaded.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def example_text() -> bytes:
    """BASELINE config 1: the reference's example/paired_dat{1,2}.fq.gz, inflated, concatenated in glob order."""
    return b"".join(gzip.open(os.path.join(GOLDEN, f)).read() for f in ("paired_dat1.fq.gz", "paired_dat2.fq.gz"))


@pytest.fixture(scope="session")
def golden() -> dict:
    return json.load(open(os.path.join(GOLDEN, "example_k31.json")))


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def hostemu():
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostemu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    L = C.CDLL(os.path.join(d, "libhostemu.so"))
    P, I64, U32, I = C.c_void_p, C.c_int64, C.c_uint32, C.c_int
    sig = {
        "emu_pack_reads": (I64, [P, P, P, I64, I, I, I, P, P, P]),
        "emu_partition": (I64, [P, I64, P, P, I64, I, I, U32]),
        "emu_partition_fetch": (None, [P, P]),
        "emu_count_records": (I64, [P, P, I64, I, I, U32, P]),
        "emu_count_fetch": (None, [P, P, P]),
        "emu_bins_bruteforce": (I64, [P, U32, I, I, U32, P]),
        "emu_fork_filter": (I64, [P, P, P, I64, I, I]),
        "emu_fork_fetch": (None, [P, P, P, P]),
        "emu_sorted_filter": (I64, [P, P, P, I64, I, I, C.c_double, I]),
        "emu_check_kmer_at": (I64, [P, I64, I]),
        "emu_table_hashes": (None, [C.c_uint64, C.c_uint64, P]),
        "emu_revcomp64": (C.c_uint64, [C.c_uint64, I]),
        "emu_revcomp128": (None, [C.c_uint64, C.c_uint64, I, P, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


@pytest.fixture(scope="session")
def rfxlib():
    """libreflexiv_cuda.so, built in tree if missing (nvcc cross-compiles without a GPU)."""
    from reflexiv_b200 import _lib
    if not os.path.exists(_lib.lib_path()):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "reflexiv_b200", "csrc"), "../libreflexiv_cuda.so"], check=True)
    return _lib.load_library()


class _Hooks:
    """tests/hooks/librfx_testhooks.so: copies of a context's packed reads / super-k-mer records (test-only)."""

    def __init__(self):
        import ctypes as C
        d = os.path.join(ROOT, "tests", "hooks")
        subprocess.run(["make", "-s", "-C", d], check=True)
        self.C = C
        self.L = C.CDLL(os.path.join(d, "librfx_testhooks.so"))
        P = C.c_void_p
        self.L.rfx_debug_reads.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), P, P, P]
        self.L.rfx_debug_records.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), P, P]

    def reads(self, ctx):
        C = self.C
        n, w = C.c_uint64(), C.c_uint64()
        assert self.L.rfx_debug_reads(ctx._ctx, C.byref(n), C.byref(w), None, None, None) == 0
        lens = np.empty(n.value, dtype=np.uint32)
        woff = np.empty(n.value, dtype=np.uint64)
        words = np.empty(w.value, dtype=np.uint64)
        assert self.L.rfx_debug_reads(ctx._ctx, C.byref(n), C.byref(w), lens.ctypes.data, woff.ctypes.data, words.ctypes.data) == 0
        return lens, woff, words

    def records(self, ctx):
        C = self.C
        n, nb = C.c_uint64(), C.c_uint32()
        assert self.L.rfx_debug_records(ctx._ctx, C.byref(n), C.byref(nb), None, None) == 0
        offs = np.empty(nb.value + 1, dtype=np.uint64)
        recs = np.empty((n.value, ctx.record_bytes() // 8), dtype=np.uint64)
        assert self.L.rfx_debug_records(ctx._ctx, C.byref(n), C.byref(nb), offs.ctypes.data, recs.ctypes.data) == 0
        return offs, recs


@pytest.fixture(scope="session")
def hooks():
    return _Hooks()


def make_reads(seed: int, genome_len: int, n_pairs: int, read_len: int = 100, err: float = 0.0, frag: int = 250) -> np.ndarray:
    """Small synthetic FASTQ through the library's host generator."""
    from workload import synth
    g = synth.genome(genome_len, seed)
    return synth.fastq(g, n_pairs, read_len=read_len, frag_len=frag, error_rate=err, seed_reads=seed + 1, seed_errors=seed + 2)
