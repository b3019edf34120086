"""Compile the two C++ selection helpers of the reference FROM WHERE THEY LIE (never copied):
/root/reference/src/skDER/skDERsum.cpp and skDERcore.cpp -> oracle/_ref/{skDERsum,skDERcore}.
They are the exact downstream consumers of the edge list (reference src/skDER/skder.py:78-80,141-143),
used by tests to check "same representatives".  The skani binary itself (Rust, un-vendored, unpinned)
cannot be built here; see DESIGN.md."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/skDER"
OUT = os.path.join(HERE, "_ref")


def build():
    """Returns {name: path} for the helpers that exist (built now or earlier)."""
    os.makedirs(OUT, exist_ok=True)
    res = {}
    for name in ("skDERsum", "skDERcore"):
        src, dst = os.path.join(REF, name + ".cpp"), os.path.join(OUT, name)
        if os.path.exists(src) and (not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", dst, src])
        if os.path.exists(dst):
            res[name] = dst
    return res


if __name__ == "__main__":
    print(build())
