"""CPU: the C-ABI library builds/loads and exports every symbol include/nanocaller_b200.h declares; with no
GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nanocaller_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from nanocaller_b200.host import capi
    lib = capi.load_library()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libnanocaller_b200.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names
    assert lib.nc_abi_version() == 1


def test_meta_layout_matches_header():
    from nanocaller_b200.host import capi
    assert capi.META_DTYPE.itemsize == 40
    assert capi.META_DTYPE.fields["sample_depth"][1] == 36
    assert ctypes.sizeof(capi.NcSnpParams) == 48


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nanocaller_b200.host import capi
    with pytest.raises(capi.NcError):
        capi.Context(0)
