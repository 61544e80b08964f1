"""Build the C-ABI CUDA library in-tree with nvcc (sm_100a only).

The product is ``dgll_b200/libdgll_b200.so`` — a plain shared library with
``extern "C"`` entry points (see ``include/dgll_b200.h``), no torch linkage.
It is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdgll_b200.so")
OBJ_DIR = os.path.join(HERE, "_obj")
SOURCES = [
    "runtime.cu", "spmm.cu", "spmm_rows.cu", "gather.cu", "gemm_simt.cu", "gemm_tcgen05.cu", "gemm_tf32.cu", "gemm.cu",
    "transpose.cu", "legacy.cu", "gat.cu", "binspmm.cu", "sampler.cu", "peer.cu", "block.cu", "layerwise.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest(path, extra=()):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    for e in extra:
        with open(e, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(verbose=False, force=False):
    """Compile every .cu that changed and relink. Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in sorted(os.listdir(CSRC)) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "dgll_b200.h"))
    nvcc = _nvcc()
    objs, dirty = [], False
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        dg = _digest(sp, headers)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read() == dg):
            continue
        dirty = True
        cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT), src, stamp, dg))
    for p, src, stamp, dg in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out.decode(errors="replace")))
        with open(stamp, "w") as f:
            f.write(dg)
    if dirty or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s" % r.stdout.decode(errors="replace"))
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
