"""L.Convolution2D / L.DilatedConvolution2D / L.EmbedID as used by net.py:12-17,34-44 and
WaveNet/modules.py:13-22,127-141 (Chainer v4 argument order)."""
import numpy

from . import functions as F
from .initializers import LeCunNormal
from .link import Link
from .variable import Parameter


def _pair(x):
    return x if isinstance(x, (tuple, list)) else (x, x)


class Convolution2D(Link):
    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, nobias=False,
                 initialW=None, initial_bias=None, dilate=1):
        super(Convolution2D, self).__init__()
        self.ksize, self.stride, self.pad, self.dilate = ksize, _pair(stride), _pair(pad), _pair(dilate)
        self.out_channels = out_channels
        with self.init_scope():
            self.W = Parameter(LeCunNormal())
            self.b = Parameter(0, (out_channels,))
        if in_channels is not None:
            self._initialize_params(in_channels)

    def _initialize_params(self, in_channels):
        kh, kw = _pair(self.ksize)
        self.W.initialize((self.out_channels, in_channels, kh, kw))

    def __call__(self, x):
        if self.W.array is None:
            self._initialize_params(x.shape[1])
        return F.convolution_2d(x, self.W, self.b, self.stride, self.pad, self.dilate)


class DilatedConvolution2D(Convolution2D):
    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, dilate=1,
                 nobias=False, initialW=None, initial_bias=None):
        super(DilatedConvolution2D, self).__init__(in_channels, out_channels, ksize, stride, pad,
                                                   nobias, initialW, initial_bias, dilate)


class EmbedID(Link):
    def __init__(self, in_size, out_size, initialW=None, ignore_label=None):
        super(EmbedID, self).__init__()
        with self.init_scope():
            self.W = Parameter(None, None)
            self.W.array = numpy.random.normal(0, 1, (in_size, out_size)).astype(numpy.float32)

    def __call__(self, x):
        idx = x.array if hasattr(x, "array") else x
        return F.Variable(self.W.array[idx])
