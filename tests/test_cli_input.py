"""Host-side input logic of the `reflexiv` driver (csrc/reflexiv_main.cpp: stream_inputs), exercised without a GPU: the
driver source is compiled into a small harness that concatenates what the consumer callback receives."""
import gzip
import os
import subprocess

import pytest

from conftest import ROOT

HARNESS = r'''
#define main reflexiv_real_main
#include "%s"
#undef main
int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "counts")) {   // counts <file> <k> <minc> <maxc>: parse_counts, rows as "w0 [w1] count"
        std::string text;
        FILE* f = fopen(argv[2], "rb");
        char buf[1 << 16]; size_t got;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
        fclose(f);
        std::vector<uint64_t> keys; std::vector<uint32_t> counts;
        const int k = atoi(argv[3]);
        if (!parse_counts(text.data(), text.size(), k, atoi(argv[4]), atoi(argv[5]), keys, counts)) { fprintf(stderr, "malformed\n"); return 2; }
        const size_t w = k <= 31 ? 1 : k / 32 + 1;
        for (size_t i = 0; i < counts.size(); i++) {
            for (size_t j = 0; j < w; j++) printf("%%llu ", (unsigned long long)keys[i * w + j]);
            printf("%%u\n", counts[i]);
        }
        return 0;
    }
    if (argc > 1 && !strcmp(argv[1], "cuts")) {     // cuts <file> <n>: where run_sharded cuts a FASTQ text into n runs of whole records
        std::string text;
        FILE* f = fopen(argv[2], "rb");
        char buf[1 << 16]; size_t got;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
        fclose(f);
        const int n = atoi(argv[3]);
        size_t prev = 0;
        for (int r = 1; r < n; r++) {
            size_t want = text.size() / n * r;
            if (want < prev) want = prev;
            prev = record_start_near(text.data(), text.size(), want);
            printf("%%zu\n", prev);
        }
        return 0;
    }
    std::string all; size_t calls = 0;
    const size_t fail_at = argc > 2 ? (size_t)atoi(argv[2]) : (size_t)-1;
    bool ok = stream_inputs(argv[1], [&](const char* d, size_t n) { all.append(d, n); return calls++ != fail_at; });
    fwrite(all.data(), 1, all.size(), stdout);
    fprintf(stderr, "ok=%%d calls=%%zu\n", ok ? 1 : 0, calls);
    return ok ? 0 : 1;
}
'''


@pytest.fixture(scope="module")
def harness(tmp_path_factory, rfxlib):
    d = tmp_path_factory.mktemp("cli")
    src = d / "t.cpp"
    src.write_text(HARNESS % os.path.join(ROOT, "reflexiv_b200", "csrc", "reflexiv_main.cpp"))
    exe = d / "t"
    libdir = os.path.join(ROOT, "reflexiv_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), str(src), "-L" + libdir, "-lreflexiv_cuda", "-lz", "-pthread", "-Wl,-rpath," + libdir], check=True)
    return str(exe)


def test_files_reach_the_consumer_in_path_order(harness, example_text, tmp_path):
    lines = example_text.split(b"\n")
    recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines) - 1, 4)]
    ind = tmp_path / "in"
    ind.mkdir()
    (ind / "_SUCCESS").write_bytes(b"")            # Spark markers and hidden files are not input
    (ind / ".part0.fq.crc").write_bytes(b"junk")
    cuts = [0, 300, 301, 900, 1500, 1500, 2000, len(recs)]
    for j in range(len(cuts) - 1):
        blob = b"".join(recs[cuts[j]:cuts[j + 1]])
        if j % 3 == 0:
            (ind / f"part{j}.fq").write_bytes(blob)            # mapped
        elif j % 3 == 1:
            (ind / f"part{j}.fq").write_bytes(blob[:-1])       # no final newline (and one empty file): copied, newline added
        else:
            with gzip.open(ind / f"part{j}.fq.gz", "wb") as f:  # inflated by a reader thread
                f.write(blob)
    for readers in ("1", "3", "16"):
        for pattern in (str(ind / "part*"), str(ind)):
            r = subprocess.run([harness, pattern], capture_output=True, env=dict(os.environ, REFLEXIV_READERS=readers))
            assert r.returncode == 0 and r.stdout == example_text, (readers, pattern, r.stderr)
            assert b"calls=6" in r.stderr                        # one push per non-empty file
    # a consumer that rejects the third file stops the stream (the readers are joined, nothing hangs)
    r = subprocess.run([harness, str(ind / "part*"), "2"], capture_output=True, timeout=60)
    assert r.returncode == 1 and b"calls=3" in r.stderr


def test_missing_and_unreadable_inputs(harness, tmp_path):
    r = subprocess.run([harness, str(tmp_path / "none*")], capture_output=True)
    assert r.returncode == 1 and b"Input path does not exist" in r.stderr
    (tmp_path / "x.4mc").write_bytes(b"\0")
    r = subprocess.run([harness, str(tmp_path / "x.4mc")], capture_output=True)
    assert r.returncode == 1 and b"hadoop-4mc" in r.stderr
    (tmp_path / "bad.fq.gz").write_bytes(gzip.compress(b"@r\nACGT\n+\nIIII\n")[:-6])   # truncated gzip stream
    r = subprocess.run([harness, str(tmp_path / "bad.fq.gz")], capture_output=True)
    assert r.returncode == 1 and b"cannot read" in r.stderr


def test_python_mirror_reads_the_same_way(example_text, tmp_path):
    """reflexiv_b200.pipeline.iter_input_files: same order, same newline rule, look-ahead on threads."""
    from reflexiv_b200.pipeline import iter_input_files, read_input_text
    lines = example_text.split(b"\n")
    recs = [b"\n".join(lines[i:i + 4]) + b"\n" for i in range(0, len(lines) - 1, 4)]
    ind = tmp_path / "in"
    ind.mkdir()
    (ind / "_SUCCESS").write_bytes(b"")
    for j, (a, b) in enumerate([(0, 700), (700, 700), (700, 1500), (1500, len(recs))]):
        blob = b"".join(recs[a:b])
        if j % 2:
            (ind / f"part{j}.fq").write_bytes(blob[:-1])
        else:
            with gzip.open(ind / f"part{j}.fq.gz", "wb") as f:
                f.write(blob)
    for readers in (1, 4):
        chunks = list(iter_input_files(str(ind), readers))
        assert len(chunks) == 3 and b"".join(chunks) == example_text and all(c.endswith(b"\n") for c in chunks)
    assert read_input_text(str(ind / "part*")) == example_text
    with pytest.raises(FileNotFoundError):
        read_input_text(str(tmp_path / "none*"))
    # classic-Mac line ends would parse as one long line (Hadoop's reader splits at a lone "\\r"): refused, "\\r\\n" is fine
    (tmp_path / "mac.fq").write_bytes(example_text[:5000].replace(b"\n", b"\r"))
    (tmp_path / "dos.fq").write_bytes(example_text[:5000].replace(b"\n", b"\r\n"))
    from reflexiv_b200 import RfxError
    with pytest.raises(RfxError):
        read_input_text(str(tmp_path / "mac.fq"))
    assert read_input_text(str(tmp_path / "dos.fq")).count(b"\r\n") == example_text[:5000].count(b"\n")


def test_count_table_parser(harness, tmp_path):
    """-kmerc rows (KmerBinarizer, ReflexivDSMain.java:3872-3948): both row forms, the >= 10-digit clamp, coverage bounds,
    one- and two-word keys in the reference layout, and the threaded path (pieces cut at line starts) against the Python mirror."""
    import numpy as np
    from reflexiv_b200.pipeline import parse_count_csv
    rng = np.random.default_rng(5)
    def kmers(n, k):
        return ["".join("ACGT"[c] for c in rng.integers(0, 4, k)) for _ in range(n)]
    def run(path, k, minc, maxc, readers="1"):
        r = subprocess.run([harness, "counts", str(path), str(k), str(minc), str(maxc)], capture_output=True, text=True, env=dict(os.environ, REFLEXIV_READERS=readers))
        return r.returncode, [tuple(int(x) for x in line.split()) for line in r.stdout.splitlines()]
    # small table, every row form
    rows31 = kmers(6, 31)
    text = f"{rows31[0]},5\n({rows31[1]},17)\n{rows31[2]},12345678901\n\n{rows31[3]},1\r\n{rows31[4]},3000\n{rows31[5]},7"
    (tmp_path / "a.csv").write_text(text)
    rc, got = run(tmp_path / "a.csv", 31, 2, 10_000_000)
    keys, cnt = parse_count_csv(text.replace("\r", "").encode(), 31)
    keep = (cnt >= 2) & (cnt <= 10_000_000)
    assert rc == 0 and got == [(int(k[0]), int(c)) for k, c in zip(keys[keep], cnt[keep])]
    assert [c for _, c in got] == [5, 17, 3000, 7]                 # the 11-digit count clamps to 10^9 and falls above maxcov
    rc, got = run(tmp_path / "a.csv", 31, -2**31, 2**31 - 1)
    assert [c for _, c in got] == [5, 17, 1_000_000_000, 1, 3000, 7]
    # two-word keys (k = 61: 32 bases, then 29 right aligned)
    rows61 = kmers(50, 61)
    text61 = "".join(f"{s},{i + 1}\n" for i, s in enumerate(rows61))
    (tmp_path / "b.csv").write_text(text61)
    rc, got = run(tmp_path / "b.csv", 61, 1, 100)
    keys, cnt = parse_count_csv(text61.encode(), 61)
    assert rc == 0 and got == [(int(k[0]), int(k[1]), int(c)) for k, c in zip(keys, cnt)]
    # malformed rows are reported, not skipped
    (tmp_path / "c.csv").write_text(f"{rows31[0]},5\nACGT,3\n")
    assert run(tmp_path / "c.csv", 31, 1, 100)[0] == 2
    # a table large enough for the threaded path: same rows in the same order with 1 and 8 parser threads
    big = kmers(4000, 31)
    text_big = "".join(f"{big[i % 4000]},{i % 90 + 1}\n" for i in range(140_000))
    (tmp_path / "d.csv").write_text(text_big)
    assert len(text_big) > (4 << 20)
    rc1, one = run(tmp_path / "d.csv", 31, 3, 80, "1")
    rc8, eight = run(tmp_path / "d.csv", 31, 3, 80, "8")
    assert rc1 == 0 and rc8 == 0 and one == eight and len(one) == sum(1 for i in range(140_000) if 3 <= i % 90 + 1 <= 80)


def test_python_count_table_parser_row_forms():
    """pipeline.parse_count_csv (vectorised): every row form against a plain per-line restatement of KmerBinarizer."""
    import numpy as np
    from reflexiv_b200.pipeline import encode_kmer_rows, parse_count_csv
    rng = np.random.default_rng(9)
    for k in (5, 31, 33, 61):
        rows = ["".join("ACGTN"[c] for c in rng.integers(0, 5, k)) for _ in range(400)]
        parts, want_k, want_c = [], [], []
        for i, r in enumerate(rows):
            c = int(rng.choice([1, 7, 42, 999, 123456789, 1234567890, 98765432109]))
            parts.append([f"({r},{c})\n", f"{r},{c}\r\n", f"{r},{c}\n\n", f"{r},{c}\n"][i % 4])
            want_k.append(r)
            want_c.append(1_000_000_000 if c >= 10 ** 9 else c)     # >= 10 digits clamp (DSMain.java:3895-3910)
        keys, counts = parse_count_csv("".join(parts).encode()[:-1], k)      # the last row without its newline
        assert np.array_equal(keys, encode_kmer_rows(want_k, k)) and counts.tolist() == want_c
    for bad in (b"ACGT,3\n", b"ACGTACGTACGTACGTACGTACGTACGTACG;3\n", b"ACGTACGTACGTACGTACGTACGTACGTACG,3x\n"):
        with pytest.raises(ValueError):
            parse_count_csv(bad, 31)
    assert parse_count_csv(b"\n\n", 31)[1].size == 0


def test_multi_rank_input_cuts_fall_on_record_starts(harness, example_text, tmp_path):
    """`reflexiv run --gpus N` gives rank r the r-th run of every file; a run must start on a record's header line -- also when
    quality lines begin with '@' -- or the ranks would read other reads than one rank does."""
    from oracle import orc
    lines = example_text.split(b"\n")
    # make every 7th quality line start with '@' and every 11th with '+'
    for j, i in enumerate(range(3, len(lines) - 1, 4)):
        if j % 7 == 0:
            lines[i] = b"@" + lines[i][1:]
        elif j % 11 == 0:
            lines[i] = b"+" + lines[i][1:]
    txt = b"\n".join(lines)
    p = tmp_path / "tricky.fq"
    p.write_bytes(txt)
    s_all, l_all = orc.fastq_reads(txt, orc.FASTQ_RUN)
    for n in (2, 3, 8):
        cuts = [int(x) for x in subprocess.run([harness, "cuts", str(p), str(n)], capture_output=True, text=True, check=True).stdout.split()]
        assert len(cuts) == n - 1 and cuts == sorted(cuts)
        bounds = [0] + cuts + [len(txt)]
        total = 0
        for a, b in zip(bounds, bounds[1:]):
            assert a == b or txt[a:a + 1] == b"@"
            s, l = orc.fastq_reads(txt[a:b], orc.FASTQ_RUN)
            total += len(s)
        assert total == len(s_all)       # the runs hold exactly the reads of the whole file


def test_lone_cr_files_are_refused_by_the_driver(harness, example_text, tmp_path):
    (tmp_path / "mac.fq").write_bytes(example_text[:4000].replace(b"\n", b"\r"))
    r = subprocess.run([harness, str(tmp_path / "mac.fq")], capture_output=True)
    assert r.returncode != 0 and b"lone" in r.stderr
    (tmp_path / "dos.fq").write_bytes(example_text[:4000].replace(b"\n", b"\r\n"))
    r = subprocess.run([harness, str(tmp_path / "dos.fq")], capture_output=True)
    assert r.returncode == 0 and r.stdout.count(b"\r\n") == example_text[:4000].count(b"\n")
