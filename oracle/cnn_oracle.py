"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

fp32 (optionally fp64) CPU restatement, in plain PyTorch ops, of the four Keras models of the
reference.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this module.

  SNP_model            model_architect.py:36-64
  haploid_SNP_model    model_architect_SNP_haploid.py:33-53
  Indel_model          model_architect_indel.py:28-48
  haploid_Indel_model  model_architect_indels_haploid.py:29-48

Keras semantics restated: Conv2D on NHWC input with HWIO kernels, 'same' = symmetric zero padding
(all 'same' kernels are odd, stride 1), 'valid', SELU (scale 1.0507009873554805, alpha
1.6732632423543772), Flatten in (H, W, C) order, Dense = x @ K + b, softmax over the last axis,
Dropout = identity at inference.

Pinned as far as it can be here: the reference's own model classes (model_architect*.py) and workers run UNCHANGED over
oracle/shim/tensorflow, which supplies only the primitive ops; the VCF records they write with the released weights
(tests/golden/records_*.vcf.txt) are reproduced by this restatement + the record restatements
(tests/test_snp_records_golden.py, tests/test_indel_records_golden.py) — so layer wiring, concatenation / flatten order, head
order and scaling are the reference's.  The primitive ops themselves stay PARITY UNPINNED against real TensorFlow numerics
(not installable in the build container; the reference ships no golden probabilities): tolerance 1e-4 on output
probabilities, per BASELINE.json north_star, is against this float32 restatement.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def _conv(x_nchw, w, name, stride, same, dtype):
    k = _t(w[name + "/kernel"], dtype).permute(3, 2, 0, 1).contiguous()   # HWIO -> OIHW
    b = _t(w[name + "/bias"], dtype)
    pad = (k.shape[2] // 2, k.shape[3] // 2) if same else 0
    return F.selu(F.conv2d(x_nchw, k, b, stride=stride, padding=pad))


def _dense(x, w, name, dtype, act=True):
    y = x @ _t(w[name + "/kernel"], dtype) + _t(w[name + "/bias"], dtype)
    return F.selu(y) if act else y


def _trunk(x, w, dtype):
    """conv1_{1,2,3} -> concat -> conv2 -> conv3 -> flatten(H,W,C) -> fc1 (shared by all four models)."""
    x = _t(x, dtype).permute(0, 3, 1, 2).contiguous()                     # NHWC -> NCHW
    c1 = torch.cat([_conv(x, w, "conv1_1", 1, True, dtype), _conv(x, w, "conv1_2", 1, True, dtype),
                    _conv(x, w, "conv1_3", 1, True, dtype)], 1)
    c2 = _conv(c1, w, "conv2", (1, 2), False, dtype)
    c3 = _conv(c2, w, "conv3", (1, 2), False, dtype)
    flat = c3.permute(0, 2, 3, 1).reshape(c3.shape[0], -1)
    return _dense(flat, w, "fc1", dtype)


@torch.no_grad()
def snp_model(w, x, ref_onehot, dtype=torch.float32):
    """model_architect.py:36-64.  x [B,5,41,5], ref_onehot [B,4] (columns A,G,T,C as fed at
    snpCaller.py:111).  Returns (out_A, out_G, out_T, out_C, out_GT), each [B,2] numpy."""
    fc1 = _trunk(x, w, dtype)
    fa = _dense(fc1, w, "fa", dtype)
    ref = _t(ref_onehot, dtype)
    outs = []
    for j, b in enumerate("AGTC"):
        outs.append(torch.softmax(_dense(torch.cat([fa, ref[:, j:j + 1]], 1), w, b, dtype, act=False), -1))
    fc2 = _dense(fc1, w, "fc2", dtype)
    fc3 = _dense(torch.cat([fc2] + outs, 1), w, "fc3", dtype)
    gt = torch.softmax(_dense(fc3, w, "GT", dtype, act=False), -1)
    return tuple(o.numpy() for o in outs) + (gt.numpy(),)


@torch.no_grad()
def haploid_snp_model(w, x, ref_onehot, dtype=torch.float32):
    """model_architect_SNP_haploid.py:33-53: fc3 is Dense(4, selu) on [fc2, ref], then softmax."""
    fc1 = _trunk(x, w, dtype)
    fc2 = _dense(fc1, w, "fc2", dtype)
    fc3 = _dense(torch.cat([fc2, _t(ref_onehot, dtype)], 1), w, "fc3", dtype)
    return torch.softmax(fc3, -1).numpy()


@torch.no_grad()
def indel_model(w, x, dtype=torch.float32):
    """model_architect_indel.py:28-48.  x [B,15,128,2] -> [B,4] softmax."""
    fc1 = _trunk(x, w, dtype)
    fc2 = _dense(fc1, w, "fc2", dtype)
    return torch.softmax(_dense(fc2, w, "fc3", dtype, act=False), -1).numpy()


@torch.no_grad()
def haploid_indel_model(w, x, dtype=torch.float32):
    """model_architect_indels_haploid.py:29-48.  x [B,5,128,2] -> [B,1] sigmoid."""
    fc1 = _trunk(x, w, dtype)
    fc2 = _dense(fc1, w, "fc2", dtype)
    return torch.sigmoid(_dense(fc2, w, "fc3", dtype, act=False)).numpy()


def snp_probs(w, x, ref_onehot, dtype=torch.float32):
    """The only outputs the caller consumes (snpCaller.py:115): [B,4] = P(A), P(G), P(T), P(C) present."""
    a, g, t, c, _ = snp_model(w, x, ref_onehot, dtype)
    return np.stack([a[:, 1], g[:, 1], t[:, 1], c[:, 1]], 1)
