"""File-level check of tests/test_host_shim.py's CASES_HOST (options that stay in the reference's host code, other penalties) on
the real GPU: oracle/_ref/kma_gpu vs kma -t 1. usage: python tools/host_options_gpu.py"""
import os, sys, tempfile, pathlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests import test_host_shim as T
ok = True
for name in T.CASES_HOST:
    with tempfile.TemporaryDirectory() as d:
        tmp = pathlib.Path(d)
        args = T._make_case(tmp, name)
        T._run("kma", args + ["-o", "ref", "-t", "1"], tmp)
        T._run("kma_gpu", args + ["-o", "gpu", "-t", "1"], tmp)
        exts = ("res", "fsa", "aln", "frag.gz", "mat.gz")
        if "-nc" in args or "-nf" in args:
            exts = tuple(e for e in exts if not (e in ("fsa", "aln") and "-nc" in args) and not (e == "frag.gz" and "-nf" in args))
        try:
            T._compare(tmp, exts=exts, at_least=min(3, len(exts) - 1))
            print(name, "equal")
        except AssertionError as e:
            ok = False
            print(name, "DIFFERS", str(e)[:120])
print("all equal" if ok else "differences")
