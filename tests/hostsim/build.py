"""TEST INFRASTRUCTURE ONLY: compile the library's .cu sources with g++ against the CUDA
execution-model emulation in cuda_hostsim.h -> tests/hostsim/_build/libdfvo_hostsim.so.
Lets the CPU test-suite exercise kernel indexing and the network orchestration without a GPU.
The product never loads this library (see df-vo_b200/b200/native.py)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "df-vo_b200", "csrc")
OUTDIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUTDIR, "libdfvo_hostsim.so")


def sources():
    sys.path.insert(0, CSRC)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_dfvo_build", os.path.join(CSRC, "build.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.SOURCES


def build(force=False):
    os.makedirs(OUTDIR, exist_ok=True)
    srcs = sources()
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    for f in ("cuda_hostsim.h", "cuda_hostsim.cpp"):
        h.update(open(os.path.join(HERE, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "..", "include", "dfvo_b200.h"), "rb").read())
    dig = h.hexdigest()
    stamp = OUT + ".stamp"
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == dig:
        return OUT
    flags = ["-O2", "-g", "-std=c++17", "-fPIC", "-DDFVO_HOSTSIM", "-I", HERE, "-I", CSRC, "-Wno-unused-value"]
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(OUTDIR, s.replace(".cu", ".o"))
        objs.append(o)
        procs.append((s, subprocess.Popen(["g++"] + flags + ["-x", "c++", "-c", os.path.join(CSRC, s), "-o", o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    o = os.path.join(OUTDIR, "cuda_hostsim.o")
    objs.append(o)
    procs.append(("cuda_hostsim.cpp", subprocess.Popen(["g++"] + flags + ["-c", os.path.join(HERE, "cuda_hostsim.cpp"), "-o", o],
                                                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("g++ failed on %s" % s)
    subprocess.check_call(["g++", "-shared", "-o", OUT] + objs)
    with open(stamp, "w") as f:
        f.write(dig)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
