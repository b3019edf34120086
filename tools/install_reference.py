"""Install the UNMODIFIED reference (raufs/skDER) into baseline/_ref so that it travels to the GPU box:
`pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>` plus its two C++ helpers,
compiled from where they lie into baseline/_ref/bin (the reference's own setup.py does the same with g++ -o).
Runs only where /root/reference exists (the build container); elsewhere the prebuilt tree is used as is."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(force=False):
    have = os.path.exists(os.path.join(DST, "skDER", "skder.py")) and all(
        os.path.exists(os.path.join(DST, "bin", b)) for b in ("skder", "skDERsum", "skDERcore"))
    if not os.path.isdir(REF) or (have and not force):
        return DST if have else None
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:  # the build writes into the source tree; /root/reference is read-only
        src = os.path.join(tmp, "ref")
        shutil.copytree(REF, src)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps",
                               "--find-links", "/opt/wheelhouse", "--upgrade", "--target", DST, src])
    for name in ("skDERsum", "skDERcore"):
        subprocess.check_call(["g++", "-O2", "-o", os.path.join(DST, "bin", name), os.path.join(REF, "src", "skDER", name + ".cpp")])
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
