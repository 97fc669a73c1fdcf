"""ORACLE / TEST INFRASTRUCTURE ONLY — a stand-in `tensorflow` so the reference's model classes
(`nanocaller_src/model_architect*.py`) and worker code (`snpCaller.caller`, `indelCaller.indel_run`) can be imported and run
UNCHANGED in the build container, where TensorFlow cannot be installed.

What is the reference's own code when run over this shim: the layer wiring of the four models (which layer feeds which,
concatenation order, flatten position, the heads), the batching, the coverage scaling, the genotype decision and the
VCF record text.  What the shim supplies, and what therefore stays "parity unpinned" against real TensorFlow numerics:
the primitive ops — Conv2D on NHWC input with HWIO kernels ('same' = symmetric zero padding for the odd kernels used,
'valid'), Dense = x @ K + b, SELU, Flatten over (H, W, C), softmax / sigmoid over the last axis, Dropout = identity at
inference, tf.concat — computed in float32 with torch.  Weights are read with this repo's checkpoint / HDF5 readers
(nanocaller_b200/host/weights.py) and attached to the layers by attribute name, which is how TF2 object checkpoints
key them.  Never imported by the product package."""
import numpy as np
import torch

from . import keras  # noqa: F401


def concat(values, axis):
    return torch.cat([_as_tensor(v) for v in values], dim=axis)


def _as_tensor(v):
    if isinstance(v, torch.Tensor):
        return v
    return torch.as_tensor(np.asarray(v, dtype=np.float32))


class _Nn:
    @staticmethod
    def selu(x):
        return torch.nn.functional.selu(x)

    @staticmethod
    def softmax(x, axis=-1):
        return torch.softmax(x, dim=axis)


nn = _Nn()
