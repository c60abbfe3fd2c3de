"""Build libcyclegan_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m cyclegan_cat_b200.build

Static cudart, no link-time dependency on libcuda or libnccl (driver entry points are resolved
at run time), so the library also loads on a CPU-only box for the symbol/ABI tests.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(LIBDIR, "libcyclegan_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "cyclegan_b200.h")]:
        if f.endswith((".h", ".cuh")) or f == os.path.basename(path):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(verbose=False, force=False):
    os.makedirs(OBJDIR, exist_ok=True)
    objs, logs = [], []
    procs = []
    for src in _sources():
        obj = os.path.join(OBJDIR, src[:-3] + ".o")
        stamp = obj + ".sha1"
        dig = _digest(os.path.join(CSRC, src))
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, stamp, dig, p in procs:
        out, _ = p.communicate()
        logs.append(f"== {src}\n{out}")
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        else:
            with open(stamp, "w") as fh:
                fh.write(dig)
    # one log per source (overwritten when that source is recompiled); build.log is their concatenation, so it always
    # holds exactly the ptxas output (registers / spills, -Xptxas -v) of the objects that are in the library
    for src, stamp, dig, p in procs:
        with open(os.path.join(OBJDIR, src + ".log"), "w") as fh:
            fh.write(next(l for l in logs if l.startswith(f"== {src}\n")))
    with open(os.path.join(LIBDIR, "build.log"), "w") as fh:
        for f in sorted(os.listdir(OBJDIR)):
            if f.endswith(".log"):
                fh.write(open(os.path.join(OBJDIR, f)).read() + "\n")
    if failed:
        raise RuntimeError("nvcc compilation failed (see above)")
    if procs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread",
                                                      "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
