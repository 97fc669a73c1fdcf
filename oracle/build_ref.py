"""ORACLE / TEST INFRASTRUCTURE ONLY — recipe that makes `oracle/_ref/` from the reference where it lies.

The reference (WGLab/NanoCaller) is pure Python, so "compiling it from its own sources" means byte-compiling the modules of
the hot path, unchanged, from /root/reference into sourceless `.pyc` files under oracle/_ref/nanocaller_src/ (git-ignored, not
gpurun-ignored: the directory travels to the GPU box like a built .so; /root/reference itself does not exist there).  The
released weights the workers look up next to their own module file (snpCaller.py:36-40, indelCaller.py:26-31) are copied as
data.  Nothing here is product code, and no reference source text enters the repository.

    python oracle/build_ref.py            (build container only; __graft_entry__.build() calls it when /root/reference exists)

Used by: bench.py --impl reference (the CPU arm runs `snpCaller.caller` / `indelCaller.indel_run` — the reference's own worker
functions — over oracle/shim), tests that compare against the reference itself.
"""
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NC_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")

MODULES = ["__init__", "utils", "generate_SNP_pileups", "snpCaller", "model_architect", "model_architect_SNP_haploid",
           "generate_indel_pileups", "generate_indel_pileups_haploid", "indelCaller", "model_architect_indel",
           "model_architect_indels_haploid"]


def available():
    return os.path.isdir(os.path.join(REF, "nanocaller_src"))


def built():
    return os.path.exists(os.path.join(OUT, "nanocaller_src", "generate_SNP_pileups.pyc"))


def build(verbose=False):
    """-> True when oracle/_ref is usable afterwards."""
    if not available():
        return built()
    src = os.path.join(REF, "nanocaller_src")
    dst = os.path.join(OUT, "nanocaller_src")
    os.makedirs(dst, exist_ok=True)
    for m in MODULES:
        p = os.path.join(src, m + ".py")
        c = os.path.join(dst, m + ".pyc")
        if not os.path.exists(p):
            if m == "__init__":
                open(os.path.join(dst, "__init__.py"), "w").close()       # namespace marker only
                continue
            raise FileNotFoundError(p)
        if not os.path.exists(c) or os.path.getmtime(c) < os.path.getmtime(p):
            py_compile.compile(p, cfile=c, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            if verbose:
                print("compiled", p, "->", c)
    # released weights (data): every model file the two workers can name
    rd_src, rd_dst = os.path.join(src, "release_data"), os.path.join(dst, "release_data")
    for root, _, files in os.walk(rd_src):
        if "bed_files" in root:
            continue
        for f in files:
            if f.endswith((".index", ".coverage", ".h5")) or ".data-" in f:
                rel = os.path.relpath(os.path.join(root, f), rd_src)
                t = os.path.join(rd_dst, rel)
                if not os.path.exists(t) or os.path.getsize(t) != os.path.getsize(os.path.join(root, f)):
                    os.makedirs(os.path.dirname(t), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), t)
    with open(os.path.join(OUT, "README"), "w") as f:
        f.write("Byte-compiled, unmodified modules of %s/nanocaller_src (oracle/build_ref.py) + the released weights.\n"
                "Test infrastructure; never imported by nanocaller_b200/.\n" % REF)
    return built()


def activate():
    """Put the shims and the byte-compiled reference on sys.path (shims first: `import pysam` must find oracle/shim/pysam.py)."""
    shim = os.path.join(HERE, "shim")
    for p in (OUT, shim):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    os.environ["PATH"] = os.path.join(shim, "bin") + os.pathsep + os.environ.get("PATH", "")      # the `muscle` stand-in


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT available (no %s)" % REF)
