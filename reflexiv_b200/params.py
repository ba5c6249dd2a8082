"""Parameter block and command-line parsing, mirroring the reference.

* ``DefaultParam``          <- util/DefaultParam.java:54-141
* ``Parameter``             <- util/Parameter.java:68-104 (option names), :311-611 (importCommandLine)   [`reflexiv run`]
* ``ParameterOfCounter``    <- util/ParameterOfCounter.java:63-82, :205-390                              [`reflexiv counter`]
* ``split_launcher_args``   <- bin/reflexiv:209-238 (``--x [value]`` goes to spark-submit, ``-x [value]`` to Reflexiv)

Paths are relative to /root/reference/src/main/java/uni/bielefeld/cmg/reflexiv/.
"""
from __future__ import annotations

import dataclasses
import sys
from typing import List, Optional, Tuple


@dataclasses.dataclass
class DefaultParam:
    inputFqPath: Optional[str] = None
    inputKmerPath: Optional[str] = None
    outputPath: Optional[str] = None
    inputFormat: str = "4mc"          # DefaultParam.java:70
    kmerSize: int = 31                # :78
    minKmerCoverage: int = 2          # :104
    maxKmerCoverage: int = 10_000_000  # :105
    minErrorCoverage: int = 8         # :106  (4 * minKmerCoverage at construction; -cover does not update it)
    minRepeatFold: float = 1.5        # :107  (Count_<k>_sorted stage)
    kmerList: str = "23,31,41,53,67,81,95"  # :87   (its largest entry + 3 is the fork flag of that stage)
    minContig: int = 500              # :108
    bubble: bool = True               # :109
    cache: bool = False
    gzip: bool = False
    partitions: int = 0               # :114
    maximumIteration: int = 150       # :115
    minimumIteration: int = 15        # :116
    frontClip: int = 0
    endClip: int = 0
    shufflePartition: int = 200       # :123
    stitch: bool = False
    minReadSize: int = 31
    readLimit: int = 2**63 - 1

    @property
    def subKmerSize(self) -> int:
        return self.kmerSize - 1

    @property
    def kmerListInt(self) -> List[int]:
        """setKmerListArray, DefaultParam.java:155-161"""
        return [int(x) for x in self.kmerList.split(",")]


class ParseExit(Exception):
    """The reference calls System.exit(0) (help, version, bad parameters, missing input): exit code 0."""

    def __init__(self, message: str = ""):
        super().__init__(message)
        self.message = message


def split_launcher_args(args: List[str]) -> Tuple[List[str], List[str]]:
    """bin/reflexiv:209-238.  Returns (spark_opts, reflexiv_opts); args exclude the command word."""
    spark, own = [], []
    n = len(args)
    for i, a in enumerate(args):
        nxt = args[i + 1] if i + 1 < n else None
        takes_value = nxt is not None and not nxt.startswith("-")
        if a in ("--spark-conf", "--spark-param", "--spark-help", "--class"):
            if takes_value:
                spark += [a, nxt]
            continue
        if a.startswith("--"):
            spark += [a] + ([nxt] if takes_value else [])
        elif a.startswith("-"):
            own += [a] + ([nxt] if takes_value else [])
    return spark, own


class _Parser:
    # name -> takes an argument (Parameter.java:150-300: OptionBuilder.hasArg())
    OPTIONS = {}

    def __init__(self, arguments: List[str]):
        self.arguments = list(arguments)
        self.param = DefaultParam()

    def _parse(self):
        """commons-cli PosixParser.parse(options, args, stopAtNonOption=true) for single-dash long names."""
        vals, i, a = {}, 0, self.arguments
        while i < len(a):
            tok = a[i]
            if not tok.startswith("-") or tok == "-":
                break  # stopAtNonOption
            name = tok.lstrip("-")
            if name not in self.OPTIONS:
                raise ParseExit(f"Parameter settings incorrect.\nUnrecognized option: {tok}")
            if self.OPTIONS[name]:
                if i + 1 >= len(a):
                    raise ParseExit(f"Parameter settings incorrect.\nMissing argument for option: {name}")
                vals[name] = a[i + 1]
                i += 2
            else:
                vals[name] = True
                i += 1
        return vals

    @staticmethod
    def _int(v: str) -> int:
        try:
            return int(v, 0)  # Integer.decode accepts 0x.. and leading-0 octal; int(v, 0) is close enough for CLI use
        except ValueError:
            try:
                return int(v)
            except ValueError:
                raise ParseExit(f"Parameter settings incorrect.\nFor input string: \"{v}\"")


class Parameter(_Parser):
    OPTIONS = dict(fastq=True, paired=True, single=True, inter=True, fasta=True, infmt=True, reads=True, contig=True,
                   kmerc=True, outfile=True, kmer=True, klist=True, gzip=False, overlap=True, miniter=True, maxiter=True,
                   clipf=True, clipe=True, cover=True, maxcov=True, error=True, bubble=False, stitch=False, minlength=True,
                   mincontig=True, partition=True, partitionredu=True, accurate=False, sbin=True, mode=True, cache=False,
                   version=False, h=False, help=False)

    HELP = ("usage: reflexiv run [--spark-options] -fastq <input fastq> -outfile <output dir> [-kmer 31] [-cover 2] "
            "[-maxcov 10000000] [-error 8] [-clipf N] [-clipe N] [-mincontig 500] [-miniter 15] [-maxiter 150] "
            "[-partition N] [-partitionredu 200] [-kmerc <Count_k csv>] [-infmt fmt] [-bubble] [-gzip] [-cache]")

    def importCommandLine(self) -> DefaultParam:
        p, v = self.param, self._parse()
        if "help" in v or "h" in v:
            raise ParseExit(self.HELP)
        if "version" in v:
            raise ParseExit("")
        p.gzip = "gzip" in v
        if "kmer" in v:
            p.kmerSize = self._int(v["kmer"])  # the range check at Parameter.java:346 is vacuous (>=1 || <=100)
        if "partition" in v:
            p.partitions = self._need(self._int(v["partition"]) >= 0, "partition", self._int(v["partition"]))
        if "partitionredu" in v:
            p.shufflePartition = self._need(self._int(v["partitionredu"]) >= 0, "partitionredu", self._int(v["partitionredu"]))
        if "klist" in v:
            p.kmerList = v["klist"]  # Parameter.java:362-382 (the per-entry range check there is vacuous); read by the sorted stage
            try:
                p.kmerListInt
            except ValueError:
                raise ParseExit(f"Parameter settings incorrect.\nFor input string: \"{v['klist']}\"")
        if "accurate" in v:
            p.minRepeatFold = 2.0  # Parameter.java:417-420
        if "bubble" in v:
            p.bubble = False  # Parameter.java:420-422
        p.stitch = "stitch" in v
        p.cache = "cache" in v
        if "miniter" in v:
            p.minimumIteration = self._need(self._int(v["miniter"]) >= 0, "miniter", self._int(v["miniter"]))
        if "maxiter" in v:
            p.maximumIteration = self._need(self._int(v["maxiter"]) <= 100000, "maxiter", self._int(v["maxiter"]))
        if "clipf" in v:
            p.frontClip = self._need(self._int(v["clipf"]) > 0, "clipf", self._int(v["clipf"]))  # :461-468
        if "clipe" in v:
            p.endClip = self._need(self._int(v["clipe"]) > 0, "clipe", self._int(v["clipe"]))    # :470-477
        if "cover" in v:
            p.minKmerCoverage = self._need(self._int(v["cover"]) >= 0, "cover", self._int(v["cover"]))  # :479-487
        if "maxcov" in v:
            p.maxKmerCoverage = self._need(self._int(v["maxcov"]) >= 0, "maxcov", self._int(v["maxcov"]))
        if "error" in v:
            p.minErrorCoverage = self._need(self._int(v["error"]) >= 0, "error", self._int(v["error"]))
        if "minlength" in v:
            p.minReadSize = self._need(self._int(v["minlength"]) >= 0, "minlength", self._int(v["minlength"]))
        if "mincontig" in v:
            p.minContig = self._need(self._int(v["mincontig"]) >= 0, "mincontig", self._int(v["mincontig"]))
        if "infmt" in v:
            p.inputFormat = v["infmt"]
        have_input = False
        if "fastq" in v:
            p.inputFqPath, have_input = v["fastq"], True
        elif "kmerc" in v:
            p.inputKmerPath, have_input = v["kmerc"], True
        if "kmerc" in v and "fastq" in v:
            p.inputKmerPath = v["kmerc"]
        if not have_input:
            raise ParseExit(self.HELP)  # Parameter.java:565-568: prints help, exit 0
        if "outfile" in v:
            p.outputPath = v["outfile"]
        else:
            raise ParseExit("Output file not set of -outfile options")
        return p

    @staticmethod
    def _need(ok: bool, name: str, value: int) -> int:
        if not ok:
            raise ParseExit(f"Parameter settings incorrect.\nParameter {name} out of range")
        return value


class ParameterOfCounter(Parameter):
    OPTIONS = dict(fastq=True, fasta=True, infmt=True, reads=True, outfile=True, gzip=False, kmer=True, overlap=True,
                   clipf=True, clipe=True, cover=True, maxcov=True, minlength=True, partition=True, partitionredu=True,
                   cache=False, version=False, h=False, help=False)
    HELP = ("usage: reflexiv counter [--spark-options] -fastq <input fastq> -outfile <output dir> [-kmer 31] [-cover 2] "
            "[-maxcov 10000000] [-clipf N] [-clipe N] [-partition N] [-partitionredu 200] [-infmt fmt] [-gzip] [-cache]")


def main_exit(e: ParseExit) -> int:
    if e.message:
        print(e.message, file=sys.stderr if e.message.startswith("Parameter settings") else sys.stdout)
    return 0
