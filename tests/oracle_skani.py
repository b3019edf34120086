"""CPU `skani` command built on the oracle -- TEST INFRASTRUCTURE ONLY.

Lets the CPU test-suite drive the UNMODIFIED reference (`skder`, tools/ref_runner.py) end to end without a GPU, and
gives the GPU tests the byte-for-byte expected outputs of the same runs.  Argument handling is the product shim's own
parser (skder_b200/cli.py parse_args: the spellings of reference src/skDER/skder.py:16-18, 58-59, 103, 119); the
arithmetic is oracle/skani_cpu.py.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv):
    from oracle import skani_cpu
    from skder_b200 import cli

    sub, opt = cli.parse_args(argv)
    t = opt["threads"]
    if sub == "triangle":
        cli.write_atomic(opt["out"], skani_cpu.triangle_tsv(cli.read_list(opt["list"]), opt["screen"], opt["min_af"], t))
    elif sub == "dist":
        cli.write_atomic(opt["out"], skani_cpu.rect_tsv(cli.read_list(opt["rl"]), cli.read_list(opt["ql"]), opt["screen"],
                                                        opt["min_af"], t))
    elif sub == "sketch":
        os.makedirs(opt["out"], exist_ok=True)
        cli.write_atomic(os.path.join(opt["out"], "manifest.json"), json.dumps({"paths": cli.read_list(opt["list"])}))
    else:  # search: every database genome against the query; the query's own copy in the database answers too
        paths = json.load(open(os.path.join(opt["db"], "manifest.json")))["paths"]
        q = opt["positional"][0]
        tsv = skani_cpu.rect_tsv([p for p in paths if p != q], [q], opt["screen"], opt["min_af"], t)
        if q in paths:  # rect_tsv skips identical paths; a genome against its own copy is 100 / ~100 / ~100
            from oracle import oracle as O

            s = O.Sketch.from_file(q)
            r = O.pair(s, s)
            row = "%s\t%s\t%.2f\t%.2f\t%.2f\t%s\t%s\n" % (q, q, r.ani * 100, r.af_a * 100, r.af_b * 100, s.first_name, s.first_name)
            lines = tsv.splitlines(keepends=True)
            body = sorted(lines[1:] + [row], key=lambda ln: -float(ln.split("\t")[2]))
            tsv = lines[0] + "".join(body)
        cli.write_atomic(opt["out"], tsv)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
