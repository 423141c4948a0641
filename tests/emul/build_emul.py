"""Compile the product's .cu sources against the CUDA execution-model emulator
(tests/emul/cuda_emul.h) with g++.  TEST INFRASTRUCTURE: the resulting
tests/emul/_build/libgnas_emul.so lets the CPU-only test tier exercise kernel
logic; the product package never loads it."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "genome-nas_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "libgnas_emul.so")
FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-fopenmp", "-DGNAS_EMUL", "-x", "c++",
         "-include", os.path.join(HERE, "cuda_emul.h"), "-Wno-unknown-pragmas", "-Wno-attributes"]


def source_hash(srcs):
    h = hashlib.sha256()
    for root in (CSRC, HERE, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode() + fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    srcs.append(os.path.join(HERE, "cuda_emul.cpp"))
    want = source_hash(srcs)
    stamp = os.path.join(BUILD, "stamp")
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == want:
        return OUT
    objs, procs = [], []
    for src in srcs:
        obj = os.path.join(BUILD, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen(["g++", *FLAGS, "-c", src, "-o", obj], stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"g++ failed on {src}")
    subprocess.check_call(["g++", "-shared", "-fopenmp", "-o", OUT, *objs])
    with open(stamp, "w") as fh:
        fh.write(want)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
