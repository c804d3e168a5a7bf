import numpy


class LeCunNormal(object):
    def __init__(self, scale=1.0):
        self.scale = scale

    def __call__(self, array):
        fan_in = int(numpy.prod(array.shape[1:])) if array.ndim > 1 else array.shape[0]
        array[...] = numpy.random.normal(0, self.scale / numpy.sqrt(fan_in), array.shape)


class Constant(object):
    def __init__(self, v):
        self.v = v

    def __call__(self, array):
        array[...] = self.v


def _get_initializer(initializer):
    if initializer is None:
        return LeCunNormal()
    if numpy.isscalar(initializer):
        return Constant(initializer)
    if isinstance(initializer, numpy.ndarray):
        return Constant(initializer)
    return initializer
