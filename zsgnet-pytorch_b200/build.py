"""In-tree build of libzsg_b200.so (nvcc, sm_100a only).  No JIT cache: the .so ships with the repo
snapshot to the GPU box.  Usage: python zsgnet-pytorch_b200/build.py [--force] [-v] [--trace]
--trace builds the diagnostics variant (-DZSG_TRACE: clock stamps in the conv kernels) as libzsg_b200_trace.so and
leaves the product library alone; tools/trace_conv.py loads it."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libzsg_b200.so")
TRACE_LIB = os.path.join(HERE, "libzsg_b200_trace.so")
SOURCES = ["api.cu", "match_loss.cu", "elementwise.cu", "lstm.cu", "data.cu", "conv_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "zsg_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, trace=False):
    if not force and not trace and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".trace.o" if trace else ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        if trace:
            cmd.insert(1, "-DZSG_TRACE")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    out = TRACE_LIB if trace else LIB
    subprocess.check_call([nvcc, "-shared", "-o", out, *objs, "-lcudart"])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
