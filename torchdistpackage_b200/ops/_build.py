"""In-tree build of the native extension ``torchdistpackage_b200/_C.so``.

* every ``csrc/**/*.cu`` is compiled by nvcc for ``sm_100a`` only
  (``-gencode arch=compute_100a,code=sm_100a -lineinfo``) -- these units do not include torch
  headers, so a rebuild takes seconds;
* ``csrc/bindings.cpp`` (the only unit that sees torch / pybind11) is compiled by g++;
* objects live under ``build/`` (git-ignored), the ``.so`` is written next to the package so it
  travels with the tree to the GPU box.

``python -m torchdistpackage_b200.ops._build`` builds from the command line; ``build()`` is what
``__graft_entry__.build()`` calls.  nvcc cross-compiles, so this works on a host without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shlex
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent.parent
CSRC = PKG_DIR / "csrc"
REPO = PKG_DIR.parent
BUILD_DIR = REPO / "build" / "tdp_b200"
SO_PATH = PKG_DIR / "_C.so"

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
_EXTRA_DEFS = [f"-D{d}" for d in os.environ.get("TDP_NVCC_DEFS", "").split() if d]
NVCC_FLAGS = [*_EXTRA_DEFS,
    "-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _cuda_home() -> str:
    for cand in (os.environ.get("CUDA_HOME"), os.environ.get("CUDA_PATH"), "/usr/local/cuda"):
        if cand and Path(cand, "bin", "nvcc").exists():
            return cand
    raise RuntimeError("nvcc not found (set CUDA_HOME)")


def _sources():
    cu = sorted(CSRC.rglob("*.cu"))
    cpp = sorted(CSRC.rglob("*.cpp"))
    return cu, cpp


def _deps_digest() -> str:
    """Headers are few; any header change rebuilds everything."""
    h = hashlib.sha1()
    for f in sorted(list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.h")) + list(CSRC.rglob("*.inc"))):
        h.update(f.read_bytes())
    return h.hexdigest()


def _needs_build(src: Path, obj: Path, stamp: str) -> bool:
    tag = obj.with_suffix(obj.suffix + ".stamp")
    if not obj.exists() or not tag.exists():
        return True
    want = hashlib.sha1(src.read_bytes()).hexdigest() + stamp
    return tag.read_text() != want


def _mark_built(src: Path, obj: Path, stamp: str) -> None:
    tag = obj.with_suffix(obj.suffix + ".stamp")
    tag.write_text(hashlib.sha1(src.read_bytes()).hexdigest() + stamp)


def _run(cmd, log: Path):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.write_text(" ".join(shlex.quote(c) for c in cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError(f"build step failed:\n{' '.join(cmd)}\n{p.stdout[-6000:]}")
    return p.stdout


def build(force: bool = False, verbose: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension as cpe

    cuda_home = _cuda_home()
    nvcc = str(Path(cuda_home, "bin", "nvcc"))
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    cu, cpp = _sources()
    stamp = _deps_digest()
    torch_inc = cpe.include_paths()
    py_inc = sysconfig.get_paths()["include"]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)

    jobs = []
    objs = []
    for src in cu:
        obj = BUILD_DIR / (src.relative_to(CSRC).as_posix().replace("/", "_") + ".o")
        objs.append(obj)
        if force or _needs_build(src, obj, stamp):
            cmd = [nvcc, *NVCC_ARCH, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
            jobs.append((src, obj, cmd))
    for src in cpp:
        obj = BUILD_DIR / (src.relative_to(CSRC).as_posix().replace("/", "_") + ".o")
        objs.append(obj)
        if force or _needs_build(src, obj, stamp):
            cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-Wno-deprecated-declarations",
                   f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=_C",
                   "-DTORCH_API_INCLUDE_EXTENSION_H",
                   "-I", str(CSRC), "-I", str(Path(cuda_home, "include")), "-I", py_inc]
            for inc in torch_inc:
                cmd += ["-isystem", inc]
            cmd += ["-c", str(src), "-o", str(obj)]
            jobs.append((src, obj, cmd))

    def _compile(job):
        src, obj, cmd = job
        out = _run(cmd, obj.with_suffix(obj.suffix + ".log"))
        _mark_built(src, obj, stamp)
        if verbose:
            print(f"[tdp build] {src.relative_to(CSRC)}")
        return out

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(_compile, jobs))

    if jobs or not SO_PATH.exists() or force:
        torch_lib = Path(torch.__file__).parent / "lib"
        cmd = ["g++", "-shared", "-o", str(SO_PATH), *[str(o) for o in objs],
               f"-L{torch_lib}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
               "-lc10_cuda", "-ltorch_cuda",
               f"-L{Path(cuda_home, 'lib64')}", "-lcudart",
               f"-Wl,-rpath,{torch_lib}", f"-Wl,-rpath,{Path(cuda_home, 'lib64')}"]
        _run(cmd, BUILD_DIR / "link.log")
        if verbose:
            print(f"[tdp build] linked {SO_PATH}")
    return SO_PATH


def ptxas_report() -> str:
    """Concatenated ``-Xptxas -v`` output of the last build (registers / spills / smem)."""
    out = []
    for log in sorted(BUILD_DIR.glob("*.cu.o.log")):
        out.append(f"==== {log.name}\n{log.read_text()}")
    return "\n".join(out)


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
