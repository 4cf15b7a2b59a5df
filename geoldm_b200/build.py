"""In-tree nvcc build of the C-ABI library (sm_100a only).  `python -m geoldm_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libgeoldm_b200.so")
SOURCES = ["api.cu", "edge_simt.cu", "edge_tc.cu", "edge_tc16.cu", "linear.cu", "optim.cu", "pack.cu", "sampler.cu", "stability.cu", "train.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "geoldm_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = None, defines=()) -> str:
    """Compile csrc/*.cu into csrc/libgeoldm_b200.so.  nvcc cross-compiles without a GPU.
    `out` / `defines`: A/B variants (another output path, extra -D flags), loaded through GEOLDM_B200_LIB."""
    if out is not None:
        return _compile(os.path.abspath(out), verbose, list(defines))
    if not force and not _stale():
        return LIB
    return _compile(LIB, verbose, list(defines))


def _compile(lib: str, verbose: bool, defines) -> str:
    """Every .cu -> object file (in parallel, cached under csrc/build/<variant>/ by mtime), then one link step."""
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libgeoldm_b200.so")
    extra = (["-DGEOLDM_TC_PROFILE"] if os.environ.get("GEOLDM_TC_PROFILE") else []) + ["-D" + d for d in defines]
    os.makedirs(os.path.dirname(lib), exist_ok=True)
    tag = "default" if not extra else "".join(c if c.isalnum() else "_" for c in "".join(extra))
    objdir = os.path.join(CSRC, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "geoldm_b200.h"))
    hdr_t = max(os.path.getmtime(h) for h in headers)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])

    def one(src):
        obj = os.path.join(objdir, src[:-3] + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) > max(hdr_t, os.path.getmtime(os.path.join(CSRC, src))):
            return obj, ""
        proc = subprocess.run([nvcc] + flags + ["-c", "-o", obj, src], cwd=CSRC, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + proc.stdout + proc.stderr)
        return obj, proc.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(one, SOURCES))
    if verbose:
        sys.stderr.write("".join(r[1] for r in results))
    proc = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC", "-o", lib] + [r[0] for r in results],
                          cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout + proc.stderr)
    return lib


if __name__ == "__main__":
    # python -m geoldm_b200.build [--force] [-v] [--out ab/variant.so] [-DNAME[=VALUE] ...]
    out_ = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out_,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
