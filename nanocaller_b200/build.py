"""In-tree native builds (no pip, no JIT cache): the built .so files sit next to the sources so
they travel to the GPU box with the repo snapshot.

  libnanocaller_b200.so   nvcc, sm_100a only: CUDA kernels + the C-ABI (include/nanocaller_b200.h)
  libnc_synth.so          g++: synthetic world generator (test / bench infrastructure)
  libnc_bamio.so          g++ -lz: native BGZF/BAM reader into the staging arrays (include/nanocaller_b200_io.h)
  libnc_phase.so          g++: read-based phasing + haplotagging between the SNP and indel stages (include/nanocaller_b200_phase.h)

`python -m nanocaller_b200.build` builds everything that is stale.
"""
import contextlib
import fcntl
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_CUDA = os.path.join(HERE, "libnanocaller_b200.so")
LIB_SYNTH = os.path.join(HERE, "libnc_synth.so")
LIB_BAMIO = os.path.join(HERE, "libnc_bamio.so")
LIB_PHASE = os.path.join(HERE, "libnc_phase.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-O3,-fno-strict-aliasing", "-shared",
]
CUDA_SOURCES = ["nc_api.cu"]
CUDA_DEPS_EXT = (".cu", ".cuh", ".h", ".hpp")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


@contextlib.contextmanager
def _locked(target):
    """One builder at a time per target: under `torch.distributed.run` every rank imports this module at once, and a rank
    must never dlopen a half-written library.  The compile goes to a temporary file that is renamed into place."""
    with open(target + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


def _compile(cmd_before_out, target, cmd_after_out, verbose):
    tmp = "%s.tmp%d" % (target, os.getpid())
    try:
        _run(list(cmd_before_out) + ["-o", tmp] + list(cmd_after_out), verbose)
        os.replace(tmp, target)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)


def build_synth(force=False, verbose=False):
    src = os.path.join(CSRC, "synth.cpp")
    with _locked(LIB_SYNTH):
        if force or _stale(LIB_SYNTH, [src]):
            _compile(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread"], LIB_SYNTH, [src], verbose)
    return LIB_SYNTH


def build_bamio(force=False, verbose=False):
    """g++ -lz -pthread: native BGZF/BAM reader (include/nanocaller_b200_io.h)."""
    src = os.path.join(CSRC, "bamio.cpp")
    hdr = os.path.join(ROOT, "include", "nanocaller_b200_io.h")
    with _locked(LIB_BAMIO):
        if force or _stale(LIB_BAMIO, [src, hdr]):
            _compile(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "include")], LIB_BAMIO, [src, "-lz"], verbose)
    return LIB_BAMIO


def build_phase(force=False, verbose=False):
    """g++ -pthread: host-side phasing / haplotagging (include/nanocaller_b200_phase.h)."""
    src = os.path.join(CSRC, "phase.cpp")
    hdr = os.path.join(ROOT, "include", "nanocaller_b200_phase.h")
    with _locked(LIB_PHASE):
        if force or _stale(LIB_PHASE, [src, hdr]):
            _compile(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "include")], LIB_PHASE, [src], verbose)
    return LIB_PHASE


def cuda_deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(CUDA_DEPS_EXT)]
    deps.append(os.path.join(ROOT, "include", "nanocaller_b200.h"))
    return deps


def build_cuda(force=False, verbose=False, extra=()):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB_CUDA):
            return LIB_CUDA  # prebuilt library travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libnanocaller_b200.so")
    with _locked(LIB_CUDA):
        if force or _stale(LIB_CUDA, cuda_deps()):
            srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
            _compile([nvcc] + NVCC_FLAGS + list(extra) + ["-I", os.path.join(ROOT, "include"), "-I", CSRC], LIB_CUDA, srcs + ["-lcudart"], verbose)
    return LIB_CUDA


def build_all(force=False, verbose=False):
    build_synth(force, verbose)
    build_bamio(force, verbose)
    build_phase(force, verbose)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
