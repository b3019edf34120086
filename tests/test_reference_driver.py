"""The UNMODIFIED reference (`skder`: bin/skder + src/skDER/skder.py + util.py + skDERsum + skDERcore, installed into
baseline/_ref by tools/install_reference.py) run end to end on test_case (BASELINE config 1) with a stand-in `skani`
first on PATH: the oracle-backed CPU command here, the B200 shim in the `-m gpu` twin, which must write the same files."""
import os
import re
import stat
import sys

import pytest

from conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _acc(p):
    return re.search(r"GCA_\d+\.\d+", os.path.basename(p)).group(0)


@pytest.fixture(scope="module")
def runner():
    import install_reference
    import ref_runner

    install_reference.install()
    if ref_runner.reference_tree() is None:
        pytest.skip("unmodified reference not installed (needs /root/reference once; baseline/_ref travels afterwards)")
    return ref_runner


def _oracle_bin(tmp_path):
    d = tmp_path / "oracle_bin"
    d.mkdir(exist_ok=True)
    s = d / "skani"
    s.write_text("#!/bin/sh\nexec %s %s \"$@\"\n" % (sys.executable, os.path.join(ROOT, "tests", "oracle_skani.py")))
    s.chmod(s.stat().st_mode | stat.S_IXUSR)
    return str(d)


def _genome_dir(tmp_path, genomes7):
    d = tmp_path / "genomes"
    d.mkdir(exist_ok=True)
    for f in genomes7:  # the reference lists a directory (util.processInputGenomes); suffix must be a FASTA one
        dst = d / os.path.basename(f)
        if not dst.exists():
            os.symlink(f, dst)
    return str(d) + "/"


def _run(runner, genomes, out, mode, tmp_path, use_oracle, ani=99.0, clusters=True):
    env_extra = None
    if use_oracle:
        env, _ = runner.env_for()
        env_extra = {"PATH": _oracle_bin(tmp_path) + os.pathsep + env["PATH"]}
    wall, reps, outdir = runner.run_skder(genomes, str(out), mode, ani, 50.0, threads=4, clusters=clusters, env_extra=env_extra,
                                          timeout=900)
    files = {}
    for name in ("skDER_Results.txt", "skDER_Clustering.txt", "Skani_Triangle_Edge_Output.txt", "Concatenated_N50.txt",
                 "Genome_Information_for_Greedy_Clustering.sorted.txt"):
        p = os.path.join(outdir, name)
        if os.path.exists(p):
            files[name] = open(p).read()
    return reps, files


GOLD_REPS_99 = ["GCA_000463665.1", "GCA_000477715.1", "GCA_018366805.1", "GCA_900186975.1", "GCA_943912955.1"]


def test_unmodified_skder_greedy_dynamic_lowmem_with_oracle_skani(runner, genomes7, tmp_path):
    """reference run_tests.sh:9 (`skder -g DIR -o skder_results/ -c 4 -n -i 99.0`) and the other two modes."""
    gdir = _genome_dir(tmp_path, genomes7)
    gold = sorted(_acc(ln) for ln in open(os.path.join(GOLDEN, "skder_results", "skDER_Results.txt")) if ln.strip())
    assert gold == GOLD_REPS_99
    reps, files = _run(runner, gdir, tmp_path / "greedy", "greedy", tmp_path, True)
    # N50s are the reference's own computation and must equal its golden file
    want_n50 = {_acc(ln.split("\t")[0]): ln.split("\t")[1].strip() for ln in open(os.path.join(GOLDEN, "skder_results", "Concatenated_N50.txt"))}
    got_n50 = {_acc(ln.split("\t")[0]): ln.split("\t")[1].strip() for ln in files["Concatenated_N50.txt"].splitlines()}
    assert got_n50 == want_n50
    assert files["Skani_Triangle_Edge_Output.txt"].count("\n") == 22  # header + 21 pairs, as the golden file
    mine = sorted(_acc(r) for r in reps)
    # The golden run has two rows printed exactly 99.00 at its own -i 99.0 cutoff (skder_results/
    # Skani_Triangle_Edge_Output.txt rows 15 and 19): membership flips on +-0.005 pp there.  Representatives must be
    # the golden five up to the genomes those two knife-edge rows decide (tests/golden/KNIFE_EDGE.md).
    knife = {"GCA_000464495.1"}
    assert set(mine) - knife == set(gold) - knife, (mine, gold)
    assert "skDER_Clustering.txt" in files and files["skDER_Clustering.txt"].count("\n") == 8  # header + 7 genomes
    # away from the knife edge the representatives are exactly what the golden edge list selects (tests/replay.py
    # runs the same reference helpers on the golden edges)
    import replay

    g7 = os.path.join(GOLDEN, "skder_results")
    for ani in (97.0, 99.5):
        reps2, _ = _run(runner, gdir, tmp_path / ("greedy%.1f" % ani), "greedy", tmp_path, True, ani=ani, clusters=False)
        want = sorted(_acc(x) for x in replay.greedy_reps(os.path.join(g7, "Skani_Triangle_Edge_Output.txt"),
                                                          os.path.join(g7, "Concatenated_N50.txt"), ani, 50.0, str(tmp_path)))
        assert sorted(_acc(r) for r in reps2) == want, ani
    # dynamic and low_mem_greedy run through their own reference code paths (skDERcore; sketch + search loop + dist)
    reps_d, files_d = _run(runner, gdir, tmp_path / "dynamic", "dynamic", tmp_path, True)
    assert 1 <= len(reps_d) <= 7 and "skDER_Clustering.txt" in files_d
    reps_l, files_l = _run(runner, gdir, tmp_path / "lowmem", "low_mem_greedy", tmp_path, True)
    assert 1 <= len(reps_l) <= 7 and files_l["skDER_Clustering.txt"].count("\n") == 8
    # low_mem_greedy is greedy by N50 alone: its representatives cover every genome at the cutoffs
    assert set(_acc(r) for r in reps_l) >= {"GCA_000463665.1", "GCA_018366805.1"} or len(reps_l) >= 3


@pytest.mark.gpu
def test_unmodified_skder_with_the_b200_shim_writes_the_same_files(runner, genomes7, tmp_path):
    """Same three runs with skder_b200/bin/skani first on PATH: every file the reference writes downstream of `skani`
    is byte-identical to the oracle-skani run (paths are the same symlinked directory)."""
    gdir = _genome_dir(tmp_path, genomes7)
    for mode in ("greedy", "dynamic", "low_mem_greedy"):
        reps_o, files_o = _run(runner, gdir, tmp_path / ("o_" + mode), mode, tmp_path, True)
        try:
            reps_g, files_g = _run(runner, gdir, tmp_path / ("g_" + mode), mode, tmp_path, False)
        finally:
            if mode == "low_mem_greedy":
                from skder_b200 import daemon

                daemon.stop_for(str(tmp_path / ("g_" + mode) / "skDER_iterative_greedy_workspace" / "skani_sketch_all.db"))
        assert reps_g == reps_o, mode
        assert files_g.keys() == files_o.keys()
        for k in files_o:
            assert files_g[k] == files_o[k], (mode, k)
