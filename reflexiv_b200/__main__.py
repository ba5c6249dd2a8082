"""`python -m reflexiv_b200 <run|counter|sort> [--spark-opts] [-reflexiv-opts]`: the launcher and entry points of the
reference in one place.

* bin/reflexiv:209-267    command word -> main class, `--x [v]` (Spark) split from `-x [v]` (Reflexiv) options
* main/Main.java:57-79, main/MainOfCounter.java:58-80    parse, then call the pipeline
* util/InfoDumper.java    "Reflexiv HH:mm:ss message" progress lines

Same behaviour as the C++ driver (csrc/reflexiv_main.cpp); `sort` is the one addition (the Count_<k>_sorted stage of the
multi-k workflows on its own, Pipelines.reflexivLeftAndRightSortingPipe).  Exit code 0 on option errors, like the reference.
"""
from __future__ import annotations

import sys
import time

from .params import Parameter, ParameterOfCounter, ParseExit, main_exit, split_launcher_args

COMMANDS = ("run", "counter", "sort")


def info(msg: str) -> None:
    print(f"Reflexiv {time.strftime('%H:%M:%S')} {msg}", flush=True)


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(Parameter.HELP)
        return 1
    cmd, rest = argv[0], argv[1:]
    if cmd not in COMMANDS:
        print(f"reflexiv: command '{cmd}' is outside the GPU path (supported: {', '.join(COMMANDS)})", file=sys.stderr)
        print(Parameter.HELP)
        return 1
    _spark, own = split_launcher_args(rest)  # spark-submit options are accepted and unused
    info({"run": "Reflexiv main initiating ... ", "counter": "Reflexiv counter initiating ... ", "sort": "Reflexiv k-mer sorting initiating ... "}[cmd])
    info("interpreting parameters.")
    try:
        param = (ParameterOfCounter if cmd == "counter" else Parameter)(own).importCommandLine()
        if cmd == "sort" and param.inputKmerPath is None:
            raise ParseExit(Parameter.HELP)
    except ParseExit as e:
        return main_exit(e)
    from .pipeline import Pipelines  # binds the CUDA library: after the option errors, as the reference parses before it starts Spark
    from ._lib import RfxError
    info("Initiating CUDA context ...")
    pipes = Pipelines(param)
    try:
        st = {"run": pipes.reflexivDSMainPipe, "counter": pipes.reflexivDSCounterPipe, "sort": pipes.reflexivLeftAndRightSortingPipe}[cmd]()
    except (RfxError, FileNotFoundError, FileExistsError) as e:
        print(f"reflexiv: {e}", file=sys.stderr)
        return 1
    info(f"done: {st['n_reads']} reads, {st['n_instances']} k-mers, {st['n_rows']} rows, {st['n_contigs']} contigs")
    return 0


if __name__ == "__main__":
    sys.exit(main())
