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
