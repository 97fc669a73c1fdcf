"""GPU: layer-by-layer parity of the tensor-core trunk (conv1+conv2 -> c2, conv3 -> c3) against the fp32
oracle, decoding the fp16 hi/lo activation images the kernels exchange through HBM."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.golden_util import golden_chunk, load_case

pytestmark = pytest.mark.gpu


def _trunk(x, n, stage, nbytes):
    from nanocaller_b200.host import capi, snp_pileups
    ctx = snp_pileups.context(0)
    lib = capi.load_library()
    lib.nc_debug_tc_trunk.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    lib.nc_debug_tc_trunk.restype = ctypes.c_int
    raw = np.zeros(nbytes, np.uint8)
    x = np.ascontiguousarray(x, np.float32)
    rc = lib.nc_debug_tc_trunk(ctx._h, x.ctypes.data, n, 0, stage, raw.ctypes.data, raw.size)
    assert rc == 0, lib.nc_last_error(ctx._h)
    return raw


def _oracle_acts(w, x):
    def conv(t, name, stride, same):
        k = torch.as_tensor(w[name + "/kernel"]).permute(3, 2, 0, 1).contiguous()
        b = torch.as_tensor(w[name + "/bias"])
        pad = (k.shape[2] // 2, k.shape[3] // 2) if same else 0
        return F.selu(F.conv2d(t, k, b, stride=stride, padding=pad))
    t = torch.as_tensor(x).permute(0, 3, 1, 2).contiguous()
    c1 = torch.cat([conv(t, "conv1_1", 1, True), conv(t, "conv1_2", 1, True), conv(t, "conv1_3", 1, True)], 1)
    c2 = conv(c1, "conv2", (1, 2), False)
    c3 = conv(c2, "conv3", (1, 2), False)
    return c2.permute(0, 2, 3, 1).numpy(), c3.permute(0, 2, 3, 1).numpy()      # NHWC


def test_trunk_stages_match_oracle():
    from nanocaller_b200.host import snp_pileups, weights as W
    from oracle import snp_oracle
    tensors, meta = W.load_model("snp", "ONT-HG002")
    ctx = snp_pileups.context(0)
    ctx.load_snp_weights(W.pack_snp_blob(tensors, False), meta["train_coverage"], False)
    rs, dct, chunks, bed, g = load_case("ont_diploid")
    w0 = golden_chunk(g, 0)
    n = 301
    x = snp_oracle.scale_counts(w0["mat"][:n], meta["train_coverage"], coverage=float(w0["depth"]))
    want_c2, want_c3 = _oracle_acts(tensors, x)

    groups = (n + 2) // 3                                       # c2 in HBM = TB's smem image: [group][part][parity][kg][site % 3][40 rows][8]
    raw = _trunk(x, n, 1, groups * 30720).view(np.float16).reshape(groups, 2, 2, 4, 3, 40, 8).astype(np.float32)
    raw = np.transpose(raw, (0, 4, 1, 2, 3, 5, 6)).reshape(groups * 3, 2, 2, 4, 40, 8)[:n]
    v = raw[:, 0] + raw[:, 1]                                   # [n, parity, kg, row, 8]
    got_c2 = np.zeros((n, 4, 20, 32), np.float32)
    for h in range(4):
        for w_ in range(20):
            got_c2[:, h, w_, :] = v[:, w_ & 1, :, h * 10 + (w_ >> 1), :].reshape(n, 32)
    err2 = np.abs(got_c2 - want_c2).max()
    assert err2 < 2e-4, err2

    tiles = (n + 127) // 128                # c3 in HBM: [tile][position 27][part][site % 128][8 chunks of 8 channels], chunk j stored at j ^ (site % 8)
    raw = _trunk(x, n, 2, tiles * 884736).view(np.float16).reshape(tiles, 27, 2, 128, 8, 8).astype(np.float32)
    unsw = np.empty_like(raw)
    for r in range(8):
        unsw[:, :, :, r::8, :, :] = raw[:, :, :, r::8, np.arange(8) ^ r, :]
    v = (unsw[:, :, 0] + unsw[:, :, 1]).reshape(tiles, 27, 128, 64)          # [tile, pos, site, ch]
    got_c3 = np.transpose(v, (0, 2, 1, 3)).reshape(tiles * 128, 27, 64)[:n].reshape(n, 3, 9, 64)
    err3 = np.abs(got_c3 - want_c3).max()
    assert err3 < 2e-4, err3
