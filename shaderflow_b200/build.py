"""
Builds libsfb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU):

    python -m shaderflow_b200.build [--force] [--verbose]

One object per translation unit under csrc/ (rebuilt only when a source or header is newer), linked
into shaderflow_b200/libsfb200.so. The .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PACKAGE = Path(__file__).resolve().parent
CSRC = PACKAGE/"csrc"
OBJ = CSRC/"build"
LIBRARY = PACKAGE/"libsfb200.so"
UNITS = ("core.cu", "audio.cu", "render.cu", "render_screen.cu", "render_frame.cu", "render_lanes.cu", "visualizer_tiled_1.cu", "visualizer_tiled_2.cu", "visualizer_tiled_3.cu", "visualizer_tiled_4.cu", "visualizer_rows.cu", "piano.cu", "pipe.cu", "sink.cu", "ingest.cu", "jit.cu", "video.cu", "flac.cu")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    for candidate in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if candidate and Path(candidate).exists():
            return candidate
    raise RuntimeError("nvcc not found; libsfb200.so cannot be built")


def stale(target: Path, sources: list[Path]) -> bool:
    return (not target.exists()) or any(s.stat().st_mtime > target.stat().st_mtime for s in sources)


def build(force: bool = False, verbose: bool = False) -> Path:
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [PACKAGE.parent/"include"/"sfb200.h"]
    OBJ.mkdir(exist_ok=True)
    jobs = []
    for unit in UNITS:
        src, obj = CSRC/unit, OBJ/(Path(unit).stem + ".o")
        if force or stale(obj, [src, *headers]):
            cmd = [nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", str(src), "-o", str(obj)]
            jobs.append((unit, cmd))

    def run(job):
        unit, cmd = job
        done = subprocess.run(cmd, capture_output=True, text=True)
        return unit, done

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1) or 1) as pool:
        for unit, done in pool.map(run, jobs):
            if verbose or done.returncode:
                sys.stderr.write(f"--- {unit}\n{done.stdout}{done.stderr}\n")
            if done.returncode:
                raise RuntimeError(f"nvcc failed on {unit}")

    objects = [OBJ/(Path(u).stem + ".o") for u in UNITS]
    if force or jobs or stale(LIBRARY, objects):
        link = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                "-o", str(LIBRARY), *map(str, objects), "-lpthread", "-lrt", "-ldl"]
        done = subprocess.run(link, capture_output=True, text=True)
        if done.returncode:
            sys.stderr.write(done.stdout + done.stderr)
            raise RuntimeError("linking libsfb200.so failed")
    return LIBRARY


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
