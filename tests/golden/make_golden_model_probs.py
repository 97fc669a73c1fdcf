"""Golden probabilities of the four reference model classes (model_architect*.py, UNMODIFIED) run over oracle/shim/tensorflow
(primitive ops only) with the released weights, on seeded inputs shaped like the callers' tensors.
Writes tests/golden/model_probs.npz.  Build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REL = "/root/reference/nanocaller_src/release_data/"


def inputs(seed=11, n=48):
    rng = np.random.RandomState(seed)
    x = np.zeros((n, 5, 41, 5), np.float32)
    x[:, 1:, :, :4] = rng.randint(-40, 41, (n, 4, 41, 4)) * np.float32(1.37)
    x[:, 0, :, :4] = np.eye(4, dtype=np.float32)[rng.randint(0, 4, (n, 41))]
    x[:, 1:, :, 4] = (rng.rand(n, 4, 41) < 0.25)
    ref = np.eye(4, dtype=np.float32)[rng.randint(0, 4, n)]
    xi = (rng.rand(n, 15, 128, 2).astype(np.float32) - 0.3) * (rng.rand(n, 15, 128, 2) < 0.4)
    return x, ref, xi.astype(np.float32)


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
    sys.path.insert(0, "/root/reference")
    from nanocaller_src.model_architect import SNP_model  # (reference, unchanged)
    from nanocaller_src.model_architect_SNP_haploid import haploid_SNP_model
    from nanocaller_src.model_architect_indel import Indel_model
    from nanocaller_src.model_architect_indels_haploid import haploid_Indel_model
    x, ref, xi = inputs()
    out = {"seed": np.array(11), "n": np.array(len(x))}
    m = SNP_model()
    m.load_weights(REL + "ONT_models/SNPs/HG002_guppy4.2.2_giab-4.2.1/model-100").expect_partial()
    a, g, t, c, gt = m([x, ref[:, 0:1].astype(np.float16), ref[:, 1:2].astype(np.float16), ref[:, 2:3].astype(np.float16), ref[:, 3:4].astype(np.float16)])
    out.update(snp_A=a, snp_G=g, snp_T=t, snp_C=c, snp_GT=gt)
    h = haploid_SNP_model()
    h.load_weights(REL + "haploid_models/SNPs/CHM13/model.24-0.9985.h5")
    out["snp_haploid"] = h([x, ref])
    i = Indel_model()
    i.load_weights(REL + "ONT_models/indels/HG002_guppy4.2_giab-4.2.1/model-100").expect_partial()
    out["indel"] = i(xi)
    hi = haploid_Indel_model()
    hi.load_weights(REL + "haploid_models/indels/CHM13/model.19-0.9811.h5")
    out["indel_haploid"] = hi(xi[:, 10:15])
    np.savez_compressed(os.path.join(HERE, "model_probs.npz"), **{k: np.asarray(v) for k, v in out.items()})
    for k, v in out.items():
        print(k, np.asarray(v).shape)


if __name__ == "__main__":
    main()
