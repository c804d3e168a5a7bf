import numpy


def _a(x):
    return x.array if isinstance(x, Variable) else x


class Variable(object):
    """Forward-only stand-in for chainer.Variable."""
    __array_priority__ = 200

    def __init__(self, data=None, **kwargs):
        self.array = _a(data)
        self.requires_grad = True

    @property
    def data(self):
        return self.array

    @data.setter
    def data(self, v):
        self.array = v

    @property
    def shape(self):
        return self.array.shape

    @property
    def ndim(self):
        return self.array.ndim

    @property
    def dtype(self):
        return self.array.dtype

    def __len__(self):
        return len(self.array)

    def __getitem__(self, idx):
        return Variable(self.array[idx])

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return Variable(self.array.reshape(shape))

    def __add__(self, o):
        return Variable(self.array + _a(o))

    __radd__ = __add__
    __iadd__ = __add__          # `h += ...` on a Variable makes a new Variable (no in-place)

    def __sub__(self, o):
        return Variable(self.array - _a(o))

    def __rsub__(self, o):
        return Variable(_a(o) - self.array)

    def __mul__(self, o):
        return Variable(self.array * _a(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return Variable(self.array / _a(o))

    def __neg__(self):
        return Variable(-self.array)

    def __pow__(self, p):
        return Variable(self.array ** p)


class Parameter(Variable):
    def __init__(self, initializer=None, shape=None, name=None):
        super(Parameter, self).__init__(None)
        self.initializer = initializer
        self.name = name
        if isinstance(initializer, numpy.ndarray):
            self.array = initializer
        elif shape is not None:
            self.initialize(shape)

    def initialize(self, shape):
        arr = numpy.empty(shape, dtype=numpy.float32)
        if callable(self.initializer):
            self.initializer(arr)
        else:
            arr[...] = 0 if self.initializer is None else self.initializer
        self.array = arr
