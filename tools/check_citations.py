"""Check that every `File.cpp:line[-line]` citation of the reference in the repo's headers, sources and docs names a file
that exists under /root/reference/pastar and has that many lines.  Run in the build container (the reference is not on
the GPU box):  python tools/check_citations.py"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pastar"
PAT = re.compile(r"\b((?:[A-Za-z_]+/)*[A-Za-z_0-9]+\.(?:cpp|hpp|h)):(\d+)(?:-(\d+))?")
OWN = {"pastar_gpu.h", "pastar_host.hpp", "pastar_main.cpp", "pg_host_weights.cpp", "boundary_test.cpp", "host_cpu_test.cpp",
       "ref_driver.cpp", "pastar_oracle.h"}


def ref_files():
    idx = {}
    for d, _, fs in os.walk(REF):
        for f in fs:
            idx.setdefault(f, []).append(os.path.join(d, f))
    return idx


def main():
    if not os.path.isdir(REF):
        print("reference not present: nothing checked")
        return 0
    idx = ref_files()
    lines = {}
    bad = checked = 0
    for d, dirs, fs in os.walk(ROOT):
        dirs[:] = [x for x in dirs if x not in (".git", "gpurun_out", "build", "_ref", "__pycache__", "superseded", "baseline")]
        for f in fs:
            if not f.endswith((".h", ".hpp", ".cpp", ".cu", ".cuh", ".c", ".py", ".md")) or f in ("SURVEY.md", "VERDICT.md", "ADVICE.md", "PAPERS.md", "SNIPPETS.md", "BASELINE.md"):
                continue
            path = os.path.join(d, f)
            for ln, text in enumerate(open(path, errors="replace"), 1):
                for m in PAT.finditer(text):
                    base = os.path.basename(m.group(1))
                    if base in OWN or base.startswith("pg_") or base.startswith("test_"):
                        continue
                    hi = int(m.group(3) or m.group(2))
                    checked += 1
                    if base not in idx:
                        print("%s:%d: cites %s, not a reference file" % (os.path.relpath(path, ROOT), ln, m.group(0)))
                        bad += 1
                        continue
                    n = max(lines.setdefault(p, sum(1 for _ in open(p, errors="replace"))) for p in idx[base])
                    if hi > n:
                        print("%s:%d: cites %s but the file has %d lines" % (os.path.relpath(path, ROOT), ln, m.group(0), n))
                        bad += 1
    print("%d citations checked, %d bad" % (checked, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
