"""The pastar CLI keeps the reference's flags, exit codes and stdout format (msa_options.cpp:24-159, msa_pastar_main.cpp:56-193)."""
import os
import re
import subprocess

import pytest

from conftest import CASES, KNOWN_OPT, ROOT, has_gpu, weighted_sp_score

BIN = os.path.join(ROOT, "mpi_pastar_msa_b200", "bin", "pastar")


def run(args, **kw):
    return subprocess.run([BIN] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, **kw)


def write_fasta(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">Sequence %d\n%s\n" % (i + 1, s))


def test_version_help_and_errors(tmp_path):
    assert os.path.exists(BIN), "python -m mpi_pastar_msa_b200.build"
    r = run(["-v"])
    assert r.returncode == 0 and r.stdout.decode() == "msa_pastar, version 1.0\n"
    r = run([])
    assert r.returncode == 1 and "[OPTIONS] file.fasta" in r.stdout.decode() and "--hash_type" in r.stdout.decode()
    r = run(["-h", "x.fasta"])
    assert r.returncode == 1 and "Usage" in r.stdout.decode()
    r = run([str(tmp_path / "missing.fasta")])
    assert r.returncode == 1 and "is not a regular file." in r.stdout.decode()
    fa = tmp_path / "a.fasta"
    write_fasta(str(fa), CASES["PF08184"])
    r = run(["-y", "BOGUS", str(fa)])
    assert r.returncode == 1 and "Invalid argument" in r.stderr.decode()
    r = run(["--frobnicate", str(fa)])
    assert r.returncode == 1 and "Invalid argument" in r.stderr.decode()


def test_option_forms_of_program_options(tmp_path):
    """msa_options.cpp:30-75 parses with Boost.ProgramOptions' default style: `--opt=v`, `--opt v`, `-o v`, `-ov`, options after
    the positional file, and long options shortened to an unambiguous prefix.  A run that gets past the parser either
    solves the input (GPU box) or ends with -1 at the first device call (no GPU); a parse error ends with 1."""
    fa = tmp_path / "a.fasta"
    write_fasta(str(fa), CASES["PF08184"])
    for args in (["--threads=2"], ["--threads", "2"], ["-t2"], ["-t", "2"], ["--hash_type=FSUM"], ["-yPSUM"], ["-s3"],
                 ["--hash_shift=3"], ["--thr", "2"], ["--hash_t=PZORDER", "--hash_s", "3"], ["--memory_debug"]):
        for argv in (args + [str(fa)], [str(fa)] + args):
            r = run(argv)
            assert r.returncode in (0, 255), (argv, r.returncode, r.stderr.decode())
            assert "Invalid argument" not in r.stderr.decode()
    for args, msg in ((["--hash", "3"], "ambiguous"), (["--batchx", "3"], "unrecognised option"), (["-t"], "Invalid argument"),
                      (["-t", "abc"], "Invalid argument"), (["--hash_type=fsum"], "Invalid argument")):
        r = run(args + [str(fa)])
        assert r.returncode == 1 and msg in r.stderr.decode(), (args, r.returncode, r.stderr.decode())
    assert run(["--ver"]).stdout.decode() == "msa_pastar, version 1.0\n"


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(tmp_path):
    fa = tmp_path / "a.fasta"
    write_fasta(str(fa), CASES["PF08184"])
    r = run([str(fa)])
    assert r.returncode == 255 and "Running fatal error" in r.stderr.decode()  # -1, as the reference on an exception
    assert "no CUDA device" in r.stderr.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [("PF08184", []), ("test", ["-t", "4", "-s", "3"]), ("test2", ["-y", "FSUM", "--batch", "64"]),
                                        ("kinase", ["--table_capacity", "134217728"])])
def test_output_format_and_score(tmp_path, name, args):
    import mpi_pastar_msa_b200 as m
    seqs = CASES[name]
    fa = tmp_path / (name + ".fasta")
    write_fasta(str(fa), seqs)
    r = run(args + [str(fa)])
    out = r.stdout.decode()
    assert r.returncode == 0, r.stderr.decode()
    lines = out.split("\n")
    assert lines[0].startswith("Starting pairwise alignments... done!")
    assert re.match(r"Phase 1 - init heuristic: \d\d:\d\d\.\d\d\d s$", lines[1])
    assert lines[2] == "Performing search with Parallel A-Star."
    assert re.match(r"Running PAStar with: \d+ threads \(1 machines with \d+ threads each\),(Full-Zorder|Full-Sum) hash, \d+ shift\.$", lines[3])
    assert re.match(r"Phase 2: PA-Star running time: \d\d:\d\d\.\d\d\d s$", lines[4])
    fin = "(" + " ".join(str(len(s)) for s in seqs) + ")"
    g = KNOWN_OPT[name]
    assert lines[5] == "Final Score: %s\tg - %d (h - 0 f - %d)" % (fin, g, g)  # Node.cpp:41-47 / Coord.cpp:29-40
    assert re.match(r"Phase 3 - backtrace: ", lines[6])
    assert re.match(r"Similarity: \d+\.\d\d%$", lines[7])
    assert lines[8] == ""
    rows = lines[9:9 + len(seqs)]  # not a tty: one unwrapped block (backtrace.cpp:20-35)
    assert weighted_sp_score(seqs, m.host_weights(seqs).astype("int32"), rows) == g
    i = 9 + len(seqs)
    assert lines[i] == "Total nodes count:"
    assert re.match(r"tid 0\tOpenList:\d+\tClosedList:\d+\tReopen:\d+\tTotal: \d+$", lines[i + 1])
    assert re.match(r"Sum\tOpenList:\d+\tClosedList:\d+\tReopen:\d+\tTotal: \d+$", lines[i + 2])


@pytest.mark.gpu
def test_metrics_json_sidecar(tmp_path):
    """--metrics_json: the counters the reference gathers on rank 0 (PAStarSyncData.cpp:13-116) as a JSON object."""
    import json
    fa, js = tmp_path / "k.fasta", tmp_path / "m.json"
    write_fasta(str(fa), CASES["fam5x60"])
    r = run(["--metrics_json", str(js), str(fa)])
    assert r.returncode == 0, r.stderr.decode()
    d = json.load(open(js))
    assert d["finished"] == 1 and d["gpus"] == 1 and len(d["partitions"]) == 1
    assert d["total"]["expansions"] == d["partitions"][0]["expansions"] > 0
    assert "Final Score:" in r.stdout.decode() and ("g - %d " % d["g"]) in r.stdout.decode()
