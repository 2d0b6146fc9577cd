"""Builds libagb200.so (all CUDA sources, sm_100a) in-tree. nvcc cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libagb200.so")
SOURCES = ["capi.cu", "tables.cu", "patterns.cu", "resnet.cu", "selfplay.cu", "solver.cu", "solver_plain.cu", "partition.cu", "dataset_api.cu"]
# host-only sources, compiled by g++ with the reference's floating-point flags (no FMA contraction, libm overloads as in the reference)
HOST_SOURCES = ["openings.cpp", "config_json.cpp"]
HOST_FLAGS = ["-O2", "-std=c++17", "-msse2", "-fPIC", "-I/usr/local/cuda/include"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v", "-rdc=false",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    hdr = os.path.join(HERE, "..", "include", "agb200.h")
    return os.path.getmtime(hdr) > t


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        extra = ["-fmad=false"] if src == "selfplay.cu" else []  # tree arithmetic follows the reference op by op (no FMA contraction)
        if src == "resnet.cu" and os.environ.get("AGB_NET_FLAGS"):
            extra = os.environ["AGB_NET_FLAGS"].split()  # experiments with the network kernel's code generation
        if src in ("solver.cu", "solver_plain.cu") and os.environ.get("AGB_SOLVER_FLAGS"):
            extra = os.environ["AGB_SOLVER_FLAGS"].split()  # experiments with the solver kernel's code generation
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src in HOST_SOURCES:
        obj = os.path.join(objdir, src.replace(".cpp", ".o"))
        cmd = [os.environ.get("CXX", "g++")] + HOST_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"compiler failed on {src}")
        objs.append(obj)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-lz", "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
