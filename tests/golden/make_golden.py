"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container only).

  python tests/golden/make_golden.py [/root/reference]

* copies the reference's example FASTQ pair (BASELINE config 1; Apache-2.0 data, not source);
* extracts the ONLY known answer the reference documents for this path, docs/example.html:303-343
  (`reflexiv run -fastq example/paired_dat*.fq.gz -kmer 31 -cover 3` -> ">Contig-4558-0" + first 1200 bases
  and two 4619-byte part files);
* records digests the oracle produces on that input (pinned by the documented vector, cross-checked against the
  digests listed in SURVEY.md section 8c).
"""
import gzip
import hashlib
import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import orc  # noqa: E402


def main(ref="/root/reference"):
    for f in ("paired_dat1.fq.gz", "paired_dat2.fq.gz"):
        shutil.copyfile(os.path.join(ref, "example", f), os.path.join(HERE, f))
        os.chmod(os.path.join(HERE, f), 0o644)
    doc = open(os.path.join(ref, "docs", "example.html")).read()
    m = re.search(r">Contig-4558-0\n([ACGT\n]+)</pre>", doc.replace("&gt;", ">"))
    prefix = m.group(1).replace("\n", "")
    assert len(prefix) == 1200
    txt = b"".join(gzip.open(os.path.join(HERE, f)).read() for f in ("paired_dat1.fq.gz", "paired_dat2.fq.gz"))
    starts, lens = orc.fastq_reads(txt, orc.FASTQ_RUN)
    out = {
        "source": "docs/example.html:303-343 + example/paired_dat{1,2}.fq.gz of rhinempi/Reflexiv",
        "documented": {"command": "reflexiv run -fastq './example/paired_dat*.fq.gz' -kmer 31 -cover 3",
                       "contig_header": ">Contig-4558-0", "contig_length": 4558, "part_file_bytes": 4619,
                       "n_part_files": 2, "prefix_1200": prefix,
                       "prefix_1200_sha256": hashlib.sha256(prefix.encode()).hexdigest()},
        "oracle": {"n_reads": int(len(starts))},
    }
    allc = orc.count_kmers(txt, starts, lens, 31)
    out["oracle"]["n_instances"] = allc["n_instances"]
    out["oracle"]["n_distinct"] = allc["n_distinct"]
    out["oracle"]["max_count"] = int(allc["counts"].max())
    for cov in (1, 2, 3):
        c = orc.count_kmers(txt, starts, lens, 31, min_count=cov, max_count=10_000_000)
        out["oracle"][f"count_ge{cov}"] = {"rows": int(len(c["counts"])), "sum": int(c["counts"].sum()),
                                           "sha256_sorted_csv": hashlib.sha256(orc.count_table_text(c, 31).encode()).hexdigest()}
    for cov in (2, 3):
        r = orc.run_pipeline(txt, k=31, cover=cov)
        cs = orc.canonical_contig_set(r["asm"]["contigs"])
        out["oracle"][f"contigs_cover{cov}"] = {"lengths": [len(x) for x in r["asm"]["contigs"]],
                                                "canonical_sha256": [hashlib.sha256(x.encode()).hexdigest() for x in cs],
                                                "fork_stats": r["forks"]["stats"]}
    # Count_31_sorted of the example (defaults minErrorCoverage 8, minRepeatFold 1.5, k-mer list up to 95).
    # The reference documents no output of this stage: these digests pin the oracle against regressions only.
    srt = orc.sorted_rows(allc["keys_hi"], allc["keys_lo"], allc["counts"], 31, 8, 1.5, 95, 1_000_000)
    out["oracle"]["sorted_k31"] = {"rows": int(len(srt["left"])), "left_forks": int((srt["left"] == 98).sum()),
                                   "right_forks": int((srt["right"] == 98).sum()),
                                   "sha256_sorted_csv": hashlib.sha256(orc.sorted_rows_text(srt, 31).encode()).hexdigest()}
    json.dump(out, open(os.path.join(HERE, "example_k31.json"), "w"), indent=1)
    print(json.dumps(out["oracle"], indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
