"""Build libpdlb200.so (sm_100a only) in-tree with nvcc.

    python -m pdl_b200.build [--force] [--jobs N]

Each .cu under csrc/ is compiled to an object (in parallel) and linked into
pdl_b200/lib/libpdlb200.so.  -fmad=false: the reference's x86-64 build has no FMA
contraction, and bit-exact IEEE + - * / parity depends on it.
The CUDA runtime is linked statically so the library has no dependency on torch's
(or Perl's) copy of libcudart.
"""
from __future__ import annotations

import argparse
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "lib" / "obj"
LIB = HERE / "lib" / "libpdlb200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xfatbin=-compress-all",   # 3x smaller .so: the in-tree library travels to the GPU box with every snapshot
]


def source_id() -> str:
    """Identity of what the library is built FROM: sha256 over every file under csrc/, include/pdlb200.h and the
    compiler flags.  It is compiled into the library (pdlb200_build_id) and compared at load time, so a prebuilt
    .so that no longer matches the sources next to it (an mtime-preserving copy, a forgotten rebuild) is refused
    instead of silently tested."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "pdlb200.h"]
    for f in files:
        h.update(f.name.encode() + b"\0" + f.read_bytes() + b"\0")
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()[:16]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(src: Path, seen=None) -> set:
    """Transitive closure of the quoted #includes of one source file."""
    seen = set() if seen is None else seen
    for inc in _INC.findall(src.read_text()):
        f = (src.parent / inc).resolve()
        if f.exists() and f not in seen:
            seen.add(f)
            _deps(f, seen)
    return seen


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    sid = source_id()
    stamp = OBJ / "api.buildid"                      # api.cu carries the id: recompile it whenever the id moves
    todo = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        if force or _stale(obj, [src, *_deps(src)]) or (src.stem == "api" and (not stamp.exists() or stamp.read_text() != sid)):
            todo.append((src, obj))

    def compile_one(pair):
        src, obj = pair
        cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
        if src.stem == "api":
            cmd.insert(1, f'-DPDLB200_BUILD_ID="{sid}"')
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if todo:
        with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            for src, r in ex.map(compile_one, todo):
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"nvcc failed on {src.name}")
                if verbose or r.stderr.strip():
                    sys.stderr.write(f"--- {src.name}\n{r.stderr}")
    stamp.write_text(sid)
    objs = [OBJ / (s.stem + ".o") for s in sources]
    if force or todo or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB),
               *map(str, objs), "-Xlinker", "--exclude-libs,ALL"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libpdlb200.so failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.jobs, a.verbose))
