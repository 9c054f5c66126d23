"""Build libmore4d_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmore4d_sm100.so")
SOURCES = ["gemm.cu", "gemm2.cu", "attention.cu", "elementwise.cu", "conv.cu", "conv_halo.cu", "vae_elementwise.cu", "project.cu", "mpm.cu", "gsplat.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "more4d_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _obj_stale(src: str, obj: str) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "more4d_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=()) -> str:
    """Compile every stale translation unit (in parallel) and link.  `defines`: extra -D macros,
    e.g. ("M4D_DEV",) for the development build with the measured kernel variants."""
    if not force and not defines and not _stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    extra = [f"-D{d}" for d in defines]

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        if not force and not defines and not _obj_stale(path, obj):
            return obj, ""
        r = subprocess.run([nvcc, *NVCC_FLAGS, *extra, "-c", path, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    logs = [l for _, l in results if l]
    cmd = [nvcc, "-shared", "-cudart", "static", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if logs:
        with open(os.path.join(CSRC, "ptxas.log"), "a" if not force else "w") as f:
            f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                defines=tuple(a[2:] for a in sys.argv if a.startswith("-D"))))
