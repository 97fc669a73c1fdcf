"""ORACLE / TEST INFRASTRUCTURE ONLY — recipe that makes `oracle/_ref/` from the reference where it lies.

The reference (WGLab/NanoCaller) is pure Python, so "compiling it from its own sources" means byte-compiling the modules of
the hot path, unchanged, from /root/reference into marshalled code objects `<module>.refbin` under oracle/_ref/nanocaller_src/
(git-ignored, not gpurun-ignored: the directory travels to the GPU box like a built .so; /root/reference itself does not exist
there; `.pyc` files do not travel, hence the own suffix and the small loader in `activate`).  The
released weights the workers look up next to their own module file (snpCaller.py:36-40, indelCaller.py:26-31) are copied as
data.  Nothing here is product code, and no reference source text enters the repository.

    python oracle/build_ref.py            (build container only; __graft_entry__.build() calls it when /root/reference exists)

Used by: bench.py --impl reference (the CPU arm runs `snpCaller.caller` / `indelCaller.indel_run` — the reference's own worker
functions — over oracle/shim), tests that compare against the reference itself.
"""
import importlib.abc
import importlib.util
import marshal
import os
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


SUFFIX = ".refbin"


def built():
    return os.path.exists(os.path.join(OUT, "nanocaller_src", "generate_SNP_pileups" + SUFFIX))


def build(verbose=False):
    """-> True when oracle/_ref is usable afterwards."""
    if not available():
        return built()
    src = os.path.join(REF, "nanocaller_src")
    dst = os.path.join(OUT, "nanocaller_src")
    os.makedirs(dst, exist_ok=True)
    for m in MODULES:
        p = os.path.join(src, m + ".py")
        c = os.path.join(dst, m + SUFFIX)
        if not os.path.exists(p):
            if m == "__init__":
                continue                                                  # the package object is made by the loader
            raise FileNotFoundError(p)
        if not os.path.exists(c) or os.path.getmtime(c) < os.path.getmtime(p):
            with open(p, "rb") as f:
                code = compile(f.read(), "nanocaller_src/%s.py" % m, "exec", dont_inherit=True)
            with open(c + ".tmp", "wb") as f:
                f.write(("%d.%d\n" % sys.version_info[:2]).encode())      # marshal is version-bound: same image here and on the GPU box
                marshal.dump(code, f)
            os.replace(c + ".tmp", c)
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


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports `nanocaller_src` and its modules from the marshalled code objects of oracle/_ref."""
    PKG = "nanocaller_src"

    def find_spec(self, name, path=None, target=None):
        d = os.path.join(OUT, self.PKG)
        if name == self.PKG:
            return importlib.util.spec_from_loader(name, self, origin=d, is_package=True)
        if name.startswith(self.PKG + ".") and os.path.exists(os.path.join(d, name.split(".", 1)[1] + SUFFIX)):
            return importlib.util.spec_from_loader(name, self, origin=os.path.join(d, name.split(".", 1)[1] + SUFFIX))
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        d = os.path.join(OUT, self.PKG)
        if module.__name__ == self.PKG:
            module.__path__ = [d]
            module.__file__ = os.path.join(d, "__init__.py")
            return
        short = module.__name__.split(".", 1)[1]
        module.__file__ = os.path.join(d, short + ".py")          # the workers look their weights up next to __file__ (snpCaller.py:38)
        with open(os.path.join(d, short + SUFFIX), "rb") as f:
            ver = f.readline().decode().strip()
            if ver != "%d.%d" % sys.version_info[:2]:
                raise ImportError("oracle/_ref was built with Python %s" % ver)
            code = marshal.load(f)
        exec(code, module.__dict__)


def activate():
    """Make `import pysam / intervaltree / parasail / tensorflow` find the stand-ins of oracle/shim and `import nanocaller_src`
    the byte-compiled reference; put the `muscle` stand-in on PATH."""
    shim = os.path.join(HERE, "shim")
    if shim in sys.path:
        sys.path.remove(shim)
    sys.path.insert(0, shim)
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    os.environ["PATH"] = os.path.join(shim, "bin") + os.pathsep + os.environ.get("PATH", "")      # the `muscle` stand-in


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT available (no %s)" % REF)
