"""GPU: pins the tcgen05 shared-memory descriptor semantics the CNN layer programs rely on
(nc_cnn_tc.cuh): K-major no-swizzle planes, row shift through the start address, K groups through LBO,
accumulation and accumulator column offsets — against numpy."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OP = np.dtype([("a_off", "<u4"), ("a_lbo", "<u4"), ("b_off", "<u4"), ("misc", "<u4")])
ACC = 1 << 15


def _probe(a_img, b_img, ops, N, ncols):
    from nanocaller_b200.host import capi, snp_pileups
    ctx = snp_pileups.context(0)
    lib = capi.load_library()
    lib.nc_debug_umma.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p, ctypes.c_int] * 2 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.nc_debug_umma.restype = ctypes.c_int
    a = np.ascontiguousarray(a_img).view(np.uint8).ravel()
    b = np.ascontiguousarray(b_img).view(np.uint8).ravel()
    prog = np.array(ops, dtype=OP)
    out = np.zeros((128, ncols), np.float32)
    rc = lib.nc_debug_umma(ctx._h, a.ctypes.data, a.size, b.ctypes.data, b.size, prog.ctypes.data, len(prog), N, ncols, out.ctypes.data)
    assert rc == 0, lib.nc_last_error(ctx._h)
    return out


def _planes(rng, n_kg, rows):
    """A planes [kg][row][8] of small exactly-representable values."""
    return (rng.randint(-8, 9, (n_kg, rows, 8)) / 4.0).astype(np.float16)


def _btile(rng, N):
    """B tile [kg 2][n][8]; returns (image, dense [16 k][N])."""
    img = (rng.randint(-8, 9, (2, N, 8)) / 8.0).astype(np.float16)
    dense = np.transpose(img.astype(np.float32), (0, 2, 1)).reshape(16, N)
    return img, dense


def test_plain_tile_and_row_shift():
    rng = np.random.RandomState(1)
    rows = 176
    A = _planes(rng, 2, rows)
    B, Bd = _btile(rng, 32)
    for shift in (0, 5, 37):
        out = _probe(A, B, [(shift * 16, rows * 16, 0, 0)], 32, 32)
        Am = np.concatenate([A[0, shift:shift + 128], A[1, shift:shift + 128]], 1).astype(np.float32)   # [128,16]
        np.testing.assert_array_equal(out, Am @ Bd)


def test_lbo_selects_another_tap_of_the_same_plane():
    rng = np.random.RandomState(2)
    rows = 200
    A = _planes(rng, 1, rows)
    B, Bd = _btile(rng, 16)
    for lbo_rows in (1, 45):
        out = _probe(A, B, [(3 * 16, lbo_rows * 16, 0, 0)], 16, 16)
        Am = np.concatenate([A[0, 3:131], A[0, 3 + lbo_rows:131 + lbo_rows]], 1).astype(np.float32)
        np.testing.assert_array_equal(out, Am @ Bd)


def test_accumulate_and_column_offsets():
    rng = np.random.RandomState(3)
    rows = 160
    A = _planes(rng, 4, rows)
    B0, B0d = _btile(rng, 16)
    B1, B1d = _btile(rng, 16)
    Bimg = np.concatenate([B0.ravel(), B1.ravel()])
    ops = [(0, rows * 16, 0, 0), (2 * rows * 16, rows * 16, 512, ACC),            # cols 0..15: two K chunks accumulated
           (7 * 16, rows * 16, 512, 16)]                                          # cols 16..31: shifted, overwrite
    out = _probe(A, Bimg, ops, 16, 32)
    A01 = np.concatenate([A[0, :128], A[1, :128]], 1).astype(np.float32)
    A23 = np.concatenate([A[2, :128], A[3, :128]], 1).astype(np.float32)
    A01s = np.concatenate([A[0, 7:135], A[1, 7:135]], 1).astype(np.float32)
    np.testing.assert_array_equal(out[:, :16], A01 @ B0d + A23 @ B1d)
    np.testing.assert_array_equal(out[:, 16:], A01s @ B1d)


def test_n48_tile():
    rng = np.random.RandomState(4)
    A = _planes(rng, 2, 128)
    B, Bd = _btile(rng, 48)
    out = _probe(A, B, [(0, 128 * 16, 0, 0)], 48, 48)
    Am = np.concatenate([A[0], A[1]], 1).astype(np.float32)
    np.testing.assert_array_equal(out, Am @ Bd)
