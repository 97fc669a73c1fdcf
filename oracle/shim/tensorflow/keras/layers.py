"""Stand-in tensorflow.keras.layers: the primitive ops in float32 torch (see oracle/shim/tensorflow/__init__.py)."""
import torch
import torch.nn.functional as F


def _act(name):
    if name is None:
        return lambda x: x
    return {"selu": F.selu, "sigmoid": torch.sigmoid, "relu": F.relu}[name]


class Layer:
    kernel = None
    bias = None


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), activation=None, name=None, use_bias=True, padding="valid", **kw):
        self.filters, self.kernel_size, self.strides, self.padding = filters, tuple(kernel_size), tuple(strides), padding
        self.act = _act(activation)
        assert use_bias

    def __call__(self, x):                         # x NHWC, kernel HWIO
        k = self.kernel.permute(3, 2, 0, 1).contiguous()
        assert tuple(self.kernel.shape[:2]) == self.kernel_size and self.kernel.shape[3] == self.filters
        pad = (self.kernel_size[0] // 2, self.kernel_size[1] // 2) if self.padding == "same" else 0
        if self.padding == "same":
            assert self.kernel_size[0] % 2 == 1 and self.kernel_size[1] % 2 == 1 and self.strides == (1, 1)
        y = F.conv2d(x.permute(0, 3, 1, 2), k, self.bias, stride=self.strides, padding=pad)
        return self.act(y).permute(0, 2, 3, 1).contiguous()


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, use_bias=True, **kw):
        self.units, self.act = units, _act(activation)
        assert use_bias

    def __call__(self, x):
        assert self.kernel.shape[1] == self.units
        return self.act(x @ self.kernel + self.bias)


class Flatten:
    def __call__(self, x):
        return x.reshape(x.shape[0], -1)


class Dropout:
    def __init__(self, rate, **kw):
        self.rate = rate

    def __call__(self, x, training=False):
        return x


class Softmax:
    def __call__(self, x):
        return torch.softmax(x, dim=-1)
