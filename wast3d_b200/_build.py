"""In-tree build of the CUDA library (sm_100a only) and of the test oracles.

`python -m wast3d_b200._build` or `__graft_entry__.build()`.

Outputs (all git-ignored, shipped to the GPU box by gpurun):
  wast3d_b200/lib/libwast3d_b200.so   the product: hand-written sm_100a kernels + C ABI
  oracle/liboracle.so                 CPU restatement of the reference (test infrastructure)
  oracle/_ref/libwast3d_ref.so        the unmodified reference CUDA sources behind a C wrapper
                                      (only rebuilt where /root/reference exists)
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "wast3d_b200" / "csrc"
LIBDIR = ROOT / "wast3d_b200" / "lib"
OBJDIR = ROOT / "build" / "obj"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(ROOT / "include"),
]


def _newer(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(map(str, cmd)) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    return r


def build_cuda(verbose: bool = False, ptxas_info: bool = False) -> Path:
    LIBDIR.mkdir(parents=True, exist_ok=True)
    OBJDIR.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "wast3d_b200.h", Path(__file__)]
    out = LIBDIR / "libwast3d_b200.so"

    def compile_one(src: Path):
        obj = OBJDIR / (src.stem + ".o")
        if _newer(obj, [src] + headers):
            flags = list(NVCC_FLAGS)
            if ptxas_info:
                flags += ["-Xptxas", "-v"]
            r = _run([NVCC] + flags + ["-c", str(src), "-o", str(obj)])
            if verbose or ptxas_info:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if _newer(out, objs):
        _run([NVCC, "-shared", "-o", str(out)] + [str(o) for o in objs] +
             ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return out


def build_oracle() -> Path | None:
    odir = ROOT / "oracle"
    srcs = sorted(odir.glob("*.c"))
    if not srcs:
        return None
    out = odir / "liboracle.so"
    if _newer(out, srcs + sorted(odir.glob("*.h"))):
        # -ffp-contract=off: the restatement's roundings must not depend on the host compiler's
        # FMA choices; explicit fmaf() is used where the reference's SASS shows an FMA.
        _run(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
              "-o", str(out)] + [str(s) for s in srcs] + ["-lm"])
    return out


def build_ref() -> Path | None:
    """Reference CUDA sources -> oracle/_ref (only where /root/reference is mounted)."""
    script = ROOT / "oracle" / "build_ref.sh"
    out = ROOT / "oracle" / "_ref" / "libwast3d_ref.so"
    ref_root = Path(os.environ.get("WAST3D_REFERENCE_ROOT", "/root/reference"))
    if not ref_root.exists():
        return out if out.exists() else None
    if _newer(out, [script, ROOT / "oracle" / "ref_wrap.cu"]):
        _run(["bash", str(script)])
    return out


def build_all(verbose: bool = False) -> dict:
    return {"cuda": build_cuda(verbose), "oracle": build_oracle(), "ref": build_ref()}


if __name__ == "__main__":
    info = "--ptxas" in sys.argv
    print(build_cuda(verbose=True, ptxas_info=info))
    if "--all" in sys.argv:
        print(build_oracle())
        print(build_ref())
