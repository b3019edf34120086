"""Run the UNMODIFIED reference (raufs/skDER: bin/skder, src/skDER/skder.py, util.py, skDERsum, skDERcore) with the
B200 `skani` stand-in first on PATH -- the drop-in as a skDER user would deploy it.

The reference tree comes from baseline/_ref (tools/install_reference.py: pip install of /root/reference + its two
C++ helpers), or from /root/reference itself where that exists.  Modules the reference imports at start-up but
never calls on the dereplication path (Bio, seaborn, matplotlib, aiofile) are absent from this image; empty
stand-ins live in tools/refstubs.  Nothing of the reference is edited or re-stated here: `skder` parses its own
arguments, lists the genomes (util.processInputGenomes, src/skDER/util.py:342-409), computes N50s
(util.determineN50, :429-474), calls `skani` through util.runCmd (:636-652) from skder.runSkaniTriangle /
lowMemGreedyDerep / runSkaniDist (src/skDER/skder.py:10-28, 95-134, 30-63) and selects representatives with
greedyDerep / dynamicDerep / determineClusters (:136-165, 65-93, 168-277).
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tools", "refstubs")
SHIM_BIN = os.path.join(ROOT, "skder_b200", "bin")


def reference_tree():
    """(python path entry, bin directory) of the unmodified reference, or None."""
    inst = os.path.join(ROOT, "baseline", "_ref")
    if os.path.exists(os.path.join(inst, "skDER", "skder.py")) and os.path.exists(os.path.join(inst, "bin", "skDERsum")):
        return inst, os.path.join(inst, "bin")
    return None


def env_for(extra_path=()):
    tree = reference_tree()
    if tree is None:
        raise RuntimeError("unmodified reference not installed: run tools/install_reference.py where /root/reference exists")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([tree[0], STUBS, ROOT] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["PATH"] = os.pathsep.join(list(extra_path) + [SHIM_BIN, tree[1], env.get("PATH", "")])
    env["LC_ALL"] = "C"
    return env, tree


def run_skder(genome_dir_or_files, outdir, mode="greedy", ani=99.0, af=50.0, threads=None, clusters=False, extra=(),
              env_extra=None, timeout=None):
    """`skder -g ... -o outdir -d mode -i ani -f af -c threads [-n]`, unmodified.  Returns (wall seconds, list of
    representative paths, outdir with trailing slash)."""
    env, tree = env_for()
    if env_extra:
        env.update(env_extra)
    g = [genome_dir_or_files] if isinstance(genome_dir_or_files, str) else list(genome_dir_or_files)
    cmd = [sys.executable, os.path.join(tree[1], "skder"), "-g"] + g + ["-o", outdir, "-d", mode, "-i", str(ani), "-f", str(af),
                                                                          "-c", str(threads or os.cpu_count() or 1)]
    if clusters:
        cmd.append("-n")
    cmd += list(extra)
    # the reference writes Skani_Dist_Output.txt RELATIVE to its working directory (bin/skder:468): run it beside its
    # output directory, not wherever the caller happens to be (only when every path given is absolute)
    work = os.path.dirname(os.path.abspath(outdir.rstrip("/")))
    movable = os.path.isabs(outdir) and all(os.path.isabs(x) for x in g) and os.path.isdir(work)
    t0, w0 = time.perf_counter(), time.time()
    p = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout,
                       cwd=work if movable else None)
    wall = time.perf_counter() - t0
    out = outdir if outdir.endswith("/") else outdir + "/"
    run_skder.last_phases = phases_from_files(out, w0, time.time())
    res = os.path.join(out, "skDER_Results.txt")
    if p.returncode != 0 or not os.path.exists(res):
        raise RuntimeError("reference skder failed (rc %d):\n%s" % (p.returncode, p.stdout[-4000:]))
    reps = [ln.strip() for ln in open(res) if ln.strip()]
    return wall, reps, out


def phases_from_files(outdir, t_start, t_end):
    """Wall seconds of the reference's stages, from the modification times of the files each stage leaves behind (its
    own Progress.log has minute resolution): {stage: seconds}.  t_start / t_end: time.time() around the run."""
    marks = [("list_genomes", "All_Genomes_Listing.txt"), ("n50", "Concatenated_N50.txt"),
             ("skani_triangle", "Skani_Triangle_Edge_Output.txt"),
             ("skDERsum", "Genome_Information_for_Greedy_Clustering.txt"),
             ("sort", "Genome_Information_for_Greedy_Clustering.sorted.txt"), ("select_representatives", "skDER_Results.txt")]
    ph, last = {}, t_start
    for name, f in marks:
        p = os.path.join(outdir, f)
        if os.path.exists(p):
            t = os.stat(p).st_mtime
            ph[name] = max(0.0, t - last)
            last = max(last, t)
    ph["copy_representatives_and_exit"] = max(0.0, t_end - last)
    return ph


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("genomes")
    ap.add_argument("outdir")
    ap.add_argument("--mode", default="greedy")
    ap.add_argument("--ani", type=float, default=99.0)
    ap.add_argument("--af", type=float, default=50.0)
    ap.add_argument("-n", action="store_true")
    a = ap.parse_args()
    w, reps, out = run_skder(a.genomes, a.outdir, a.mode, a.ani, a.af, clusters=a.n)
    print("%.2f s, %d representatives" % (w, len(reps)))
    print(run_skder.last_phases)
