"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/reflexiv_cuda.h
declares, refuses to work without a GPU (no CPU fallback), and the host-side mirror of the reference's option
parsing behaves like Parameter.importCommandLine."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "reflexiv_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rfx_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(rfxlib):
    from reflexiv_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(rfxlib, n), f"{n} declared in include/reflexiv_cuda.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert rfxlib.rfx_version().startswith(b"reflexiv_cuda")


def test_default_params_match_reference_defaults(rfxlib):
    from reflexiv_b200 import _lib
    p = _lib.RfxParams()
    assert rfxlib.rfx_params_default(C.byref(p)) == 0
    assert p.struct_size == C.sizeof(_lib.RfxParams)
    # util/DefaultParam.java:78,104-116,123
    assert (p.kmer_size, p.min_kmer_coverage, p.max_kmer_coverage, p.min_error_coverage) == (31, 2, 10_000_000, 8)
    assert (p.min_contig, p.bubble, p.min_iter, p.max_iter, p.shuffle_partitions) == (500, 1, 15, 150, 200)


def test_no_cpu_fallback(rfxlib):
    """Without a CUDA device rfx_create must fail with a message; nothing silently runs on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import reflexiv_b200 as R
    with pytest.raises(R.RfxError) as e:
        R.ReflexivContext(R.DefaultParam())
    assert e.value.code == R._lib.RFX_E_CUDA and "CUDA" in str(e.value)


def test_bad_params_rejected(rfxlib):
    import reflexiv_b200 as R
    p = R.DefaultParam(kmerSize=64)
    with pytest.raises(R.RfxError) as e:
        R.ReflexivContext(p)
    assert e.value.code == R._lib.RFX_E_INVALID


def test_synth_generator_is_deterministic_and_well_formed(rfxlib, orc):
    from workload import synth
    g = synth.genome(5000)
    assert set(np.unique(g).tolist()) <= set(b"ACGT") and np.array_equal(g, synth.genome(5000))
    a = synth.fastq(g, 300, read_len=150, frag_len=400)
    b = synth.fastq(g, 300, read_len=150, frag_len=400)
    assert np.array_equal(a, b)
    # split generation == whole generation (per-pair hashing, no sequential state)
    lo = synth.fastq(g, 100, first_pair=0)
    hi = synth.fastq(g, 200, first_pair=100)
    half = len(a) // 2
    assert bytes(a[:half]) == bytes(lo[:len(lo) // 2]) + bytes(hi[:len(hi) // 2])
    s1, l1 = orc.fastq_reads(a, orc.FASTQ_RUN)
    s2, l2 = orc.fastq_reads(a, orc.FASTQ_COUNTER)
    assert len(s1) == 600 and np.array_equal(s1, s2) and set(l1.tolist()) == {150}
    txt = bytes(a)
    gs = bytes(g).decode()
    for st in s1[:50]:
        r = txt[int(st):int(st) + 150].decode()
        assert r in gs or orc.revcomp_str(r) in gs
    e = synth.fastq(g, 300, error_rate=0.01)
    diff = (np.frombuffer(bytes(e), np.uint8) != np.frombuffer(bytes(a), np.uint8)).mean()
    assert 0.002 < diff < 0.008  # ~1 % of the bases, which are ~48 % of the bytes


def test_launcher_and_option_parsing():
    from reflexiv_b200.params import Parameter, ParameterOfCounter, ParseExit, split_launcher_args
    spark, own = split_launcher_args(["--driver-memory", "3G", "--executor-memory", "3G", "-fastq", "./example/paired_dat*.fq.gz",
                                      "-outfile", "./example/result", "-kmer", "31", "-cover", "3", "-bubble", "-gzip"])
    assert spark == ["--driver-memory", "3G", "--executor-memory", "3G"]
    p = Parameter(own).importCommandLine()
    assert (p.kmerSize, p.minKmerCoverage, p.minErrorCoverage, p.bubble, p.gzip) == (31, 3, 8, False, True)
    assert p.inputFqPath == "./example/paired_dat*.fq.gz" and p.outputPath == "./example/result"
    # -cover does not touch minErrorCoverage (Parameter.java:479-487); -error does
    assert Parameter(["-fastq", "x", "-outfile", "o", "-error", "3"]).importCommandLine().minErrorCoverage == 3
    # clips must be > 0 when given (Parameter.java:461-477); failures exit with code 0
    for bad in (["-fastq", "x", "-outfile", "o", "-clipf", "0"], ["-outfile", "o"], ["-fastq", "x", "-nosuch"], ["-help"]):
        with pytest.raises(ParseExit):
            Parameter(bad).importCommandLine()
    # -klist / -accurate feed the Count_<k>_sorted stage (Parameter.java:362-387, 417-420; DefaultParam.java:87, 107)
    d = Parameter(["-kmerc", "x", "-outfile", "o"]).importCommandLine()
    assert (d.kmerListInt[-1], d.minRepeatFold) == (95, 1.5)
    q = Parameter(["-kmerc", "x", "-outfile", "o", "-klist", "21,33", "-accurate"]).importCommandLine()
    assert (q.kmerListInt, q.minRepeatFold) == ([21, 33], 2.0)
    with pytest.raises(ParseExit):
        Parameter(["-kmerc", "x", "-outfile", "o", "-klist", "21,zz"]).importCommandLine()
    # -kmerc next to -fastq keeps both: the table is assembled (Pipelines.java:83-84), the FASTQ is only read by -stitch (Parameter.java:571-575)
    b = Parameter(["-fastq", "y", "-kmerc", "x", "-outfile", "o", "-stitch"]).importCommandLine()
    assert (b.inputKmerPath, b.inputFqPath, b.stitch) == ("x", "y", True)
    assert Parameter(["-kmerc", "x", "-outfile", "o"]).importCommandLine().stitch is False
    with pytest.raises(ParseExit):
        ParameterOfCounter(["-fastq", "x", "-outfile", "o", "-mincontig", "5"]).importCommandLine()  # not a counter option
    c = ParameterOfCounter(["-fastq", "x", "-outfile", "o", "-kmer", "61", "-infmt", "line"]).importCommandLine()
    assert (c.kmerSize, c.inputFormat) == (61, "line")


def test_contig_and_key_formatting_helpers():
    from reflexiv_b200.pipeline import decode_keys, encode_kmer_rows, format_contig, keys_to_int, parse_count_csv
    s = format_contig("A" * 250, -4, -5, 7)
    assert s == ">Contig-250-(-4,-5)-7\n" + "A" * 100 + "\n" + "A" * 100 + "\n" + "A" * 50
    assert format_contig("C" * 200, 30, -1, 0).count("\n") == 2
    for k in (5, 31, 33, 61, 63):
        rng = np.random.default_rng(k)
        kmers = ["".join("ACGT"[i] for i in rng.integers(0, 4, k)) for _ in range(20)]
        rows = encode_kmer_rows(kmers, k)
        assert rows.shape == (20, 1 if k <= 31 else k // 32 + 1)
        assert decode_keys(rows, k) == kmers
        assert keys_to_int(rows, k) == [int("".join(format("ACGT".index(c), "02b") for c in s), 2) for s in kmers]
    keys, counts = parse_count_csv(b"ACGTA,12\n(TTTTT,3)\nGGGGG,12345678901\n", 5)
    assert counts.tolist() == [12, 3, 1_000_000_000] and decode_keys(keys, 5) == ["ACGTA", "TTTTT", "GGGGG"]


def test_python_launcher_mirrors_the_reference_exit_codes(tmp_path):
    """`python -m reflexiv_b200 <command>` (bin/reflexiv + main/Main.java): option errors print and exit 0 like the
    reference (Parameter.java:601-611), commands outside the GPU path and a missing GPU exit 1 -- never a CPU fallback."""
    import subprocess
    import sys
    from conftest import ROOT, GOLDEN
    def run(*a):
        return subprocess.run([sys.executable, "-m", "reflexiv_b200", *a], capture_output=True, text=True, cwd=ROOT)
    r = run("run", "--driver-memory", "3G", "-fastq", "x", "-nosuch")
    assert r.returncode == 0 and "Parameter settings incorrect" in r.stderr and "Reflexiv" in r.stdout
    r = run("run", "-outfile", str(tmp_path / "o"))
    assert r.returncode == 0 and "usage: reflexiv run" in r.stdout and not (tmp_path / "o").exists()
    r = run("sort", "-fastq", "x", "-outfile", str(tmp_path / "o"))        # the sorted stage reads a count table
    assert r.returncode == 0 and "usage" in r.stdout
    r = run("meta", "-fastq", "x")
    assert r.returncode == 1 and "outside the GPU path" in r.stderr
    import torch
    if not torch.cuda.is_available():
        r = run("counter", "-fastq", os.path.join(GOLDEN, "paired_dat*.fq.gz"), "-outfile", str(tmp_path / "c"), "-kmer", "31")
        assert r.returncode == 1 and "CUDA" in r.stderr and not (tmp_path / "c" / "Count_31").exists()
