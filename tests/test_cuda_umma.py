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


BLBO0 = 1 << 14


def test_b_lbo_zero_reads_one_k_group_twice():
    """LBO = 0 on the B descriptor: both K groups of the MMA read the same 8-wide weight group, so
    (a_hi | a_lo) x (w ; w) = (a_hi + a_lo) w needs the weights only once in shared memory (conv1 of the SNP trunk)."""
    rng = np.random.RandomState(5)
    rows = 140
    A = _planes(rng, 2, rows)
    for N in (32, 64, 96):
        B = (rng.randint(-8, 9, (1, N, 8)) / 8.0).astype(np.float16)           # one K group: [n][8]
        Bd = B[0].astype(np.float32).T                                         # [8][N]
        out = _probe(A, B, [(4 * 16, rows * 16, 0, BLBO0)], N, 96 if N > 64 else 64)[:, :N]
        Am = (A[0, 4:132].astype(np.float32) + A[1, 4:132].astype(np.float32))
        np.testing.assert_array_equal(out, Am @ Bd)


def test_mixed_n_accumulation_into_overlapping_columns():
    """MMAs of different N accumulate into overlapping accumulator column ranges in issue order
    (conv1: centre tap N = 96 first, then N = 64 / N = 32 taps into sub-ranges)."""
    rng = np.random.RandomState(6)
    rows = 200
    A = _planes(rng, 2, rows)
    B96 = (rng.randint(-8, 9, (1, 96, 8)) / 8.0).astype(np.float16)
    B64 = (rng.randint(-8, 9, (1, 64, 8)) / 8.0).astype(np.float16)
    B32 = (rng.randint(-8, 9, (1, 32, 8)) / 8.0).astype(np.float16)
    Bimg = np.concatenate([B96.ravel(), B64.ravel(), B32.ravel()])

    def n8(n):
        return (n // 8) << 16
    ops = [(0, rows * 16, 0, BLBO0 | n8(96)),
           (7 * 16, rows * 16, 96 * 16, BLBO0 | ACC | n8(64) | 32),           # columns 32..95
           (50 * 16, rows * 16, (96 + 64) * 16, BLBO0 | ACC | n8(32) | 32),   # columns 32..63
           (9 * 16, rows * 16, 96 * 16, BLBO0 | ACC | n8(64) | 0)]            # columns 0..63
    out = _probe(A, Bimg, ops, 96, 96)

    def am(shift):
        return A[0, shift:shift + 128].astype(np.float32) + A[1, shift:shift + 128].astype(np.float32)
    want = am(0) @ B96[0].astype(np.float32).T
    want[:, 32:96] += am(7) @ B64[0].astype(np.float32).T
    want[:, 32:64] += am(50) @ B32[0].astype(np.float32).T
    want[:, 0:64] += am(9) @ B64[0].astype(np.float32).T
    np.testing.assert_array_equal(out, want)


def test_n_override_reads_a_prefix_of_a_wider_tile():
    """conv2 / conv3: a_hi x [w_hi | w_lo] is one N = 2C MMA, a_lo x w_hi an N = C MMA on the same tile (same LBO)."""
    rng = np.random.RandomState(7)
    A = _planes(rng, 4, 128)
    B, Bd = _btile(rng, 64)
    ops = [(0, 128 * 16, 0, 0), (2 * 128 * 16, 128 * 16, 0, ACC | ((32 // 8) << 16))]
    out = _probe(A, B, ops, 64, 64)
    A01 = np.concatenate([A[0], A[1]], 1).astype(np.float32)
    A23 = np.concatenate([A[2], A[3]], 1).astype(np.float32)
    want = A01 @ Bd
    want[:, :32] += A23 @ Bd[:, :32]
    np.testing.assert_array_equal(out, want)
