"""GPU: the tensor-core indel CNN (nc_cnn_tc_indel.cuh: IA conv1 + conv2, IB conv3, IC fc1 + tail) against the fp32 oracle —
layer by layer (decoding the fp16 hi/lo images the kernels exchange through HBM) and end to end, for Indel_model
(model_architect_indel.py:28-48) and haploid_Indel_model (model_architect_indels_haploid.py:29-48)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _inputs(n, H, seed):
    """Tensors shaped like msa() output (generate_indel_pileups.py:54-71): per 5-row group, channel 1 = one-hot of the reference
    row, channel 0 = column frequencies minus that one-hot; trailing columns zero."""
    rng = np.random.default_rng(seed)
    x = np.zeros((n, H, 128, 2), np.float32)
    for g in range(H // 5):
        depth = rng.integers(2, 60, (n, 1, 1))
        cnt = rng.multinomial(1, [0.3, 0.2, 0.2, 0.2, 0.1], (n, 128)).astype(np.float32)          # reference row one-hot
        freq = rng.dirichlet([0.5] * 5, (n, 128)).astype(np.float32)
        freq = np.round(freq * depth) / np.maximum(1, np.round(freq * depth).sum(2, keepdims=True))
        ncol = rng.integers(60, 129, n)
        live = (np.arange(128)[None, :] < ncol[:, None]).astype(np.float32)[:, :, None]
        x[:, 5 * g:5 * g + 5, :, 1] = np.transpose(cnt * live, (0, 2, 1))
        x[:, 5 * g:5 * g + 5, :, 0] = np.transpose((freq.astype(np.float32) - cnt) * live, (0, 2, 1))
    return x


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


def _trunk(ctx, x, haploid, stage, nbytes):
    from nanocaller_b200.host import capi
    lib = capi.load_library()
    lib.nc_debug_tci_trunk.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
    lib.nc_debug_tci_trunk.restype = ctypes.c_int
    raw = np.zeros(nbytes, np.uint8)
    x = np.ascontiguousarray(x, np.float32)
    rc = lib.nc_debug_tci_trunk(ctx._h, x.ctypes.data, len(x), 1 if haploid else 0, stage, raw.ctypes.data, raw.size)
    assert rc == 0, lib.nc_last_error(ctx._h)
    return raw


@pytest.mark.parametrize("haploid", [False, True])
def test_indel_trunk_stages_match_oracle(haploid):
    from nanocaller_b200.host import snp_pileups, weights as W
    tensors, _ = W.load_model("indel", "haploid" if haploid else "ONT-HG002")
    ctx = snp_pileups.context(0)
    ctx.load_indel_weights(W.pack_indel_blob(tensors), haploid)
    H = 5 if haploid else 15
    n = 37
    x = _inputs(n, H, 5)
    want_c2, want_c3 = _oracle_acts(tensors, x)
    H2, H3, NS = H - 1, H - 2, (H - 2 + 3) // 4
    raw = _trunk(ctx, x, haploid, 1, n * NS * 40960).view(np.float16).reshape(n, NS, 2, 2, 4, 160, 8).astype(np.float32)
    v = raw[:, :, 0] + raw[:, :, 1]                              # [n, slab, parity, kg, row, 8]
    got_c2 = np.zeros((n, H2, 63, 32), np.float32)
    for h in range(H2):
        for w_ in range(63):
            got_c2[:, h, w_, :] = v[:, h // 4, w_ & 1, :, (h % 4) * 32 + (w_ >> 1), :].reshape(n, 32)
    err2 = np.abs(got_c2 - want_c2).max() / np.abs(want_c2).max()                 # activations reach the hundreds: relative to the layer's range
    assert err2 < 5e-6, (err2, np.abs(want_c2).max())
    for sl in range(1, NS):                                       # the row two slabs share is stored in both
        a = raw[:, sl - 1, :, :, :, 4 * 32:4 * 32 + 31]
        b = raw[:, sl, :, :, :, 0:31]
        assert np.array_equal(a, b)
    npos = H3 * 31
    raw = _trunk(ctx, x, haploid, 2, n * npos * 192).view(np.float16).reshape(n, npos, 2, 48).astype(np.float32)
    got_c3 = (raw[:, :, 0] + raw[:, :, 1]).reshape(n, H3, 31, 48)
    err3 = np.abs(got_c3 - want_c3).max() / np.abs(want_c3).max()
    assert err3 < 5e-6, (err3, np.abs(want_c3).max())


@pytest.mark.parametrize("model,haploid", [("ONT-HG002", False), ("CCS-HG002", False), ("NanoCaller1", False), ("haploid", True)])
def test_indel_model_tensor_core_matches_oracle(model, haploid):
    """Drop-in indel_model(x) / hap_indel_model(x) (indelCaller.py:85,171) on tcgen05 vs the fp32 oracle, several tiles and a ragged tail."""
    from nanocaller_b200.host import snp_pileups, weights as W
    from oracle import cnn_oracle
    tensors, _ = W.load_model("indel", model)
    ctx = snp_pileups.context(0)
    ctx.load_indel_weights(W.pack_indel_blob(tensors), haploid)
    n = 333
    x = _inputs(n, 5 if haploid else 15, 9)
    got = ctx.indel_model_forward(x, haploid=haploid, impl=0)
    want = cnn_oracle.haploid_indel_model(tensors, x) if haploid else cnn_oracle.indel_model(tensors, x)
    err = float(np.abs(got - np.asarray(want).reshape(got.shape)).max())
    assert err < 1e-4, err
    f32 = ctx.indel_model_forward(x, haploid=haploid, impl=1)
    assert float(np.abs(f32 - np.asarray(want).reshape(got.shape)).max()) < 1e-4
