"""Make the unmodified reference available where /root/reference does not exist (the GPU box): copy the files its CPU
hot path imports -- code/*.py and configs/ -- into oracle/_ref/ (git-ignored: reference sources never enter this
repository's history; NOT gpurun-ignored, so the copy travels to the box with the snapshot).  bench.py's reference arm
and cpu_baseline leg run it through oracle/ref_harness.py.  Called by __graft_entry__.build() when /root/reference is
present; a no-op otherwise."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def build(verbose=False):
    if not os.path.exists(os.path.join(SRC, "code", "mdl.py")):
        return os.path.exists(os.path.join(DST, "code", "mdl.py"))
    for sub, pat in (("code", ".py"), ("configs", ".json")):
        os.makedirs(os.path.join(DST, sub), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(SRC, sub))):
            if f.endswith(pat):
                shutil.copyfile(os.path.join(SRC, sub, f), os.path.join(DST, sub, f))
                if verbose:
                    print("copied", sub + "/" + f)
    return True


if __name__ == "__main__":
    print("reference available:", build(verbose="-v" in sys.argv))
