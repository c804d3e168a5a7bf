"""NumPy shim of the Chainer v4 primitives that the reference's net.py, utils.py and
WaveNet/modules.py call -- TEST INFRASTRUCTURE (golden-vector generation only).

Chainer is not installable in this environment, so the reference's own model code cannot run.
This package restates, forward-only and on NumPy arrays, the semantics of exactly the
`chainer.*` names those three files touch (conv = im2col + tensordot cross-correlation like
chainer.functions.connection.convolution_2d; resize_images = the float64 linspace bilinear
gather of chainer.functions.array.resize_images; and so on), so that
`oracle/make_golden.py` can import the reference's files UNMODIFIED from /root/reference and
freeze their outputs under tests/golden/.  Nothing in the product imports it.
"""
import contextlib

import numpy

from . import configuration, cuda, function_node, functions, initializers, link, links  # noqa
from . import reporter, utils, variable  # noqa
from .configuration import config, using_config  # noqa
from .link import Chain, ChainList, Link  # noqa
from .variable import Parameter, Variable  # noqa

__version__ = "4.0.0b3-shim"


class function(object):
    @staticmethod
    @contextlib.contextmanager
    def force_backprop_mode():
        yield


class training(object):
    class StandardUpdater(object):
        pass

    class ParallelUpdater(object):
        pass
