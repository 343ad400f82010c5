"""Build libggad_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m ggad_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libggad_b200.so")
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _cutlass_includes():
    """Header-only CUTLASS / CuTe tree for the tcgen05 GEMM (dense_*.cu): vendored under site-packages in this image."""
    import sysconfig
    roots = [os.environ.get("CUTLASS_DIR")] + [os.path.join(sysconfig.get_paths()["purelib"], p)
                                                for p in ("flashinfer/data/cutlass", "tilelang/3rdparty/cutlass")]
    for r in roots:
        if r and os.path.exists(os.path.join(r, "include", "cutlass", "gemm", "collective", "builders", "sm100_9xBF16_umma_builder.inl")):
            return ["-I" + os.path.join(r, "include"), "-I" + os.path.join(r, "tools", "util", "include"), "-diag-suppress", "20012"]
    raise RuntimeError("CUTLASS >= 4.x headers with the sm100 builders not found (set CUTLASS_DIR)")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ggad_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """Compile csrc/*.cu and link the shared library.  ``out`` / ``defines`` build an experimental variant
    (e.g. for A/B runs selected with the GGAD_B200_LIB environment variable)."""
    global LIB
    variant = out is not None
    if not variant and not force and not _stale():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build" if not variant else "build_" + os.path.basename(out).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)

    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
              [os.path.join(HERE, "..", "include", "ggad_b200.h"), os.path.abspath(__file__)]
    stamp = os.path.join(objdir, ".flags")
    flags_now = " ".join(sorted(defines)) + "|" + " ".join(NVCC_FLAGS)
    same_flags = os.path.exists(stamp) and open(stamp).read() == flags_now

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        # incremental: an object newer than its source and than the headers it can include is kept
        deps = [os.path.join(CSRC, src)] + [h for h in headers
                                            if src.startswith("dense") or not h.endswith("dense_cutlass.cuh")]
        if same_flags and not force and os.path.exists(obj) and all(os.path.getmtime(obj) > os.path.getmtime(d) for d in deps):
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *["-D" + d for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if src.startswith("dense_"):
            cmd += _cutlass_includes()
        if verbose:
            cmd += ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            with open(os.path.join(objdir, src + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
        return obj

    with ThreadPoolExecutor(min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    with open(stamp, "w") as f:
        f.write(flags_now)
    target = out if variant else LIB
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv, out=outs[0] if outs else None,
                        defines=defs))
