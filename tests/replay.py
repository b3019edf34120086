"""Downstream replay: edge list -> representatives with the reference's OWN selection code.

greedy : skDERsum (reference src/skDER/skDERsum.cpp, compiled into oracle/_ref by oracle/build_ref.py)
         -> `sort -k 2 -gr` -> the loop of reference src/skDER/skder.py:150-165, restated below.
dynamic: skDERcore (reference src/skDER/skDERcore.cpp), as reference src/skDER/skder.py:78-80 runs it.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def helpers():
    from oracle import build_ref

    return build_ref.build()


def greedy_reps(edge_tsv, n50_tsv, ani, af, workdir):
    h = helpers()
    gi = os.path.join(workdir, "gi.txt")
    with open(gi, "w") as f:
        subprocess.check_call([h["skDERsum"], edge_tsv, n50_tsv, str(ani), str(af)], stdout=f)
    srt = os.path.join(workdir, "gi.sorted.txt")
    with open(srt, "w") as f:
        subprocess.check_call(["sort", "-k", "2", "--parallel=2", "-gr", gi], stdout=f, env=dict(os.environ, LC_ALL="C"))
    reps, accounted = [], set()
    with open(srt) as f:  # reference skder.py:150-165
        for line in f:
            ls = line.rstrip("\n").split("\t")
            if ls[0] in accounted:
                continue
            reps.append(ls[0])
            accounted.add(ls[0])
            if len(ls) > 2 and ls[2].strip():
                for m in ls[2].split("; "):
                    accounted.add(m.strip())
    return reps


def dynamic_reps(edge_tsv, n50_tsv, ani, af, max_af_diff):
    h = helpers()
    out = subprocess.check_output([h["skDERcore"], edge_tsv, n50_tsv, str(ani), str(af), str(max_af_diff)], text=True)
    return [ln for ln in out.splitlines() if ln.strip()]
