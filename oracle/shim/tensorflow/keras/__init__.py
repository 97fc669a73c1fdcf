"""Stand-in tensorflow.keras (see oracle/shim/tensorflow/__init__.py)."""
import numpy as np
import torch

from . import layers, regularizers  # noqa: F401


class _Out(np.ndarray):
    """numpy array that also answers `.numpy()` like an EagerTensor."""

    def numpy(self):
        return np.asarray(self)


def _out(t):
    return np.asarray(t.detach().numpy()).view(_Out)


class _LoadStatus:
    def expect_partial(self):
        return self


class Model:
    def __init__(self, *a, **kw):
        self._loaded = False

    def _layers(self):
        return {k: v for k, v in self.__dict__.items() if isinstance(v, layers.Layer)}

    def load_weights(self, path):
        """TF2 object checkpoint (keys = attribute names) or Keras HDF5 (assigned in layer order), read by the repo's own readers."""
        import os
        from nanocaller_b200.host import weights as W
        kind = "indel" if "indel" in type(self).__name__.lower() else "snp"
        if str(path).endswith(".h5"):
            tensors = W.read_keras_h5(path, kind)
        else:
            assert os.path.exists(path + ".index"), path
            tensors = W.read_tf_checkpoint(path)
        for name, layer in self._layers().items():
            if name + "/kernel" in tensors:
                layer.kernel = torch.as_tensor(np.asarray(tensors[name + "/kernel"], np.float32))
                layer.bias = torch.as_tensor(np.asarray(tensors[name + "/bias"], np.float32))
        self._loaded = True
        return _LoadStatus()

    def build(self, input_shape=None):
        return None

    def __call__(self, inputs, training=False):
        if not self._loaded:                       # the reference builds the haploid model with one dummy call before loading
            return None
        with torch.no_grad():
            if isinstance(inputs, (list, tuple)):
                inputs = [torch.as_tensor(np.asarray(x, dtype=np.float32)) for x in inputs]
            else:
                inputs = torch.as_tensor(np.asarray(inputs, dtype=np.float32))
            out = self.call(inputs)
        if isinstance(out, (list, tuple)):
            return tuple(_out(o) for o in out)
        return _out(out)
