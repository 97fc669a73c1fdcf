"""Host mirror of the reference's model objects (model_architect*.py) over the CUDA library.

The callables keep the reference call signatures so worker code reads the same:
    snp_model([x, A_ref, G_ref, T_ref, C_ref])   -> (out_A, out_G, out_T, out_C, out_GT)   snpCaller.py:111
    hap_snp_model([x, ref])                      -> probs [B,4]                            snpCaller.py:183
    indel_model(x)                               -> probs [B,4]                            indelCaller.py:85
    hap_indel_model(x)                           -> probs [B,1]                            indelCaller.py:171
All arithmetic runs in libnanocaller_b200.so; `impl` 0 = tcgen05 kernels, 1 = fp32 CUDA-core kernels."""
import numpy as np

from . import snp_pileups, weights as W


def get_SNP_model(snp_model, nanocaller_src=None):
    """snpCaller.get_SNP_model (snpCaller.py:36-55): -> (tensors, train_coverage) or (None, None)."""
    tensors, meta = W.load_model("snp", snp_model, nanocaller_src)
    if tensors is None:
        return None, None
    return tensors, meta["train_coverage"]


def get_indel_model(indel_model, nanocaller_src=None):
    """indelCaller.get_indel_model (indelCaller.py:26-38)."""
    tensors, _ = W.load_model("indel", indel_model, nanocaller_src)
    return tensors


class SNP_model:
    def __init__(self, tensors, train_coverage=0.0, device=0, impl=0):
        self.ctx = snp_pileups.context(device)
        self.impl = impl
        self.train_coverage = train_coverage
        self.ctx.load_snp_weights(W.pack_snp_blob(tensors, False), train_coverage, False)

    def __call__(self, inputs):
        x, a_ref, g_ref, t_ref, c_ref = inputs
        ref = np.concatenate([np.asarray(r, np.float32).reshape(-1, 1) for r in (a_ref, g_ref, t_ref, c_ref)], 1)
        out = self.ctx.snp_model_forward(x, ref, haploid=False, impl=self.impl)
        return tuple(out[:, 2 * j:2 * j + 2] for j in range(5))


class haploid_SNP_model:
    def __init__(self, tensors, device=0, impl=0):
        self.ctx = snp_pileups.context(device)
        self.impl = impl
        self.ctx.load_snp_weights(W.pack_snp_blob(tensors, True), 30.0, True)      # hap_train_coverage, snpCaller.py:73

    def __call__(self, inputs):
        x, ref = inputs
        return self.ctx.snp_model_forward(x, ref, haploid=True, impl=self.impl)


class Indel_model:
    def __init__(self, tensors, device=0, impl=1):
        self.ctx = snp_pileups.context(device)
        self.impl = impl
        self.ctx.load_indel_weights(W.pack_indel_blob(tensors), False)

    def __call__(self, x):
        return self.ctx.indel_model_forward(x, haploid=False, impl=self.impl)


class haploid_Indel_model:
    def __init__(self, tensors, device=0, impl=1):
        self.ctx = snp_pileups.context(device)
        self.impl = impl
        self.ctx.load_indel_weights(W.pack_indel_blob(tensors), True)

    def __call__(self, x):
        return self.ctx.indel_model_forward(x, haploid=True, impl=self.impl)
