"""Builds libathenak_b200.so (CUDA kernels + C ABI + host layer) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("ATHENAK_B200_LIB") or os.path.join(HERE, "libathenak_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "athenak_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out/defines: build a tuning variant into another file (used by profiles/ sweeps)."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    # an object is reused when it is newer than its source, every header of csrc/ and include/, this file and the flag set
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [
        os.path.join(HERE, "..", "include", "athenak_b200.h"), os.path.abspath(__file__)]
    stamp = os.path.join(objdir, "flags.txt")
    flagtext = " ".join(NVCC_FLAGS + [f"-D{d}" for d in defines])
    same_flags = os.path.exists(stamp) and open(stamp).read() == flagtext
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        if (not force and same_flags and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_hdr)):
            continue
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-x", "cu", "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose:
            print(log)
    with open(stamp, "w") as f:
        f.write(flagtext)
    target = out or LIB
    subprocess.check_call([nvcc, "-shared", "-o", target, *objs, "-lcudart"])
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
