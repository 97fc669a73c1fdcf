"""CPU: the ctypes / numpy mirrors of the C-ABI structs have the layout the headers declare.  The headers are compiled
with gcc as plain C (they must stay C: no torch types, no C++ in the signatures) into a program that prints sizeof and
offsetof of every field; the result is compared with nanocaller_b200/host/capi.py and bamio.py."""
import ctypes
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STRUCTS = {
    "nanocaller_b200.h": ["NcSnpParams", "NcChunk", "NcSiteMeta", "NcTimings", "NcIndelParams", "NcIndelVariant", "NcIndelSiteMeta", "NcIndelTimings", "NcBamDeviceContig"],
    "nanocaller_b200_io.h": ["NcBamContig"],
}


def _fields(header, name):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (name, name), src, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        out += [re.sub(r"\[.*\]", "", n.strip()) for n in names.split(",")]
    return out


def _c_layout(tmp_path):
    lines = ["#include <stdio.h>", "#include <stddef.h>"] + ['#include "%s"' % h for h in STRUCTS] + ["int main(void) {"]
    for h, names in STRUCTS.items():
        for n in names:
            lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (n, n))
            for f in _fields(h, n):
                lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (n, f, n, f))
    lines += ["return 0; }"]
    src, exe = str(tmp_path / "layout.c"), str(tmp_path / "layout")
    open(src, "w").write("\n".join(lines))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
    out = {}
    for ln in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines():
        s, f, v = ln.split()
        out.setdefault(s, {})[f] = int(v)
    return out


def _py_layout(obj):
    if isinstance(obj, np.dtype):
        d = {k: v[1] for k, v in obj.fields.items()}
        d["sizeof"] = obj.itemsize
        return d
    d = {n: getattr(obj, n).offset for n, _ in obj._fields_}
    d["sizeof"] = ctypes.sizeof(obj)
    return d


def test_struct_layouts_match_the_headers(tmp_path):
    from nanocaller_b200.host import bamio, capi
    want = _c_layout(tmp_path)
    mirrors = {"NcSnpParams": capi.NcSnpParams, "NcChunk": capi.CHUNK_DTYPE, "NcSiteMeta": capi.META_DTYPE, "NcTimings": capi.NcTimings,
               "NcIndelParams": capi.NcIndelParams, "NcIndelVariant": capi.VARIANT_DTYPE, "NcIndelSiteMeta": capi.INDEL_META_DTYPE, "NcIndelTimings": capi.NcIndelTimings, "NcBamDeviceContig": capi.NcBamDeviceContig,
               "NcBamContig": bamio.NcBamContig}
    assert sorted(mirrors) == sorted(want)
    for name, m in mirrors.items():
        assert _py_layout(m) == want[name], name


def test_order_variants_keeps_the_dict_semantics_of_variants_and_extra_variants():
    """generate_indel_pileups.py:268,:274,:301-302,:309: a later hit on a key overwrites its type; extra_variants keeps the source
    column of the last imputed hit on the key, also when a phased hit overwrites the type afterwards; keys are visited in column order."""
    from nanocaller_b200.host import capi
    from nanocaller_b200.host.indel_pileups import order_variants
    v = np.array([(90, 1, 0, 100),      # imputed hit at column 100 -> key 90
                  (90, 0, 0, 0),        # phased large-window hit at column 130 -> the same key, type overwritten, read sets kept
                  (300, 1, 0, 0),
                  (200, 1, 0, 210), (200, 1, 0, 0),
                  (50, 0, 1, 0), (40, 1, 1, 50)], dtype=capi.VARIANT_DTYPE)
    got = order_variants(v)
    assert got.dtype == capi.VARIANT_DTYPE
    assert got.tolist() == [(90, 0, 0, 100), (200, 1, 0, 210), (300, 1, 0, 0), (40, 1, 1, 50), (50, 0, 1, 0)]
    assert len(order_variants(np.zeros(0, capi.VARIANT_DTYPE))) == 0
