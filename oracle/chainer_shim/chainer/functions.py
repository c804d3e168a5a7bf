"""chainer.functions subset, forward-only on NumPy (see package docstring)."""
import numpy

from .variable import Variable, _a


def _pair(x):
    return x if isinstance(x, (tuple, list)) else (x, x)


def _out(size, k, s, p, d):
    return (size + 2 * p - d * (k - 1) - 1) // s + 1


def im2col_cpu(img, kh, kw, sy, sx, ph, pw, dy=1, dx=1):
    n, c, h, w = img.shape
    out_h, out_w = _out(h, kh, sy, ph, dy), _out(w, kw, sx, pw, dx)
    img = numpy.pad(img, ((0, 0), (0, 0), (ph, ph + sy - 1), (pw, pw + sx - 1)), mode="constant")
    col = numpy.ndarray((n, c, kh, kw, out_h, out_w), dtype=img.dtype)
    for j in range(kh):
        jdy = j * dy
        j_lim = jdy + sy * out_h
        for i in range(kw):
            idx = i * dx
            i_lim = idx + sx * out_w
            col[:, :, j, i, :, :] = img[:, :, jdy:j_lim:sy, idx:i_lim:sx]
    return col


def convolution_2d(x, W, b=None, stride=1, pad=0, dilate=1):
    x, W, b = _a(x), _a(W), _a(b)
    sy, sx = _pair(stride)
    ph, pw = _pair(pad)
    dy, dx = _pair(dilate)
    kh, kw = W.shape[2:]
    col = im2col_cpu(x, kh, kw, sy, sx, ph, pw, dy, dx)
    y = numpy.tensordot(col, W, ((1, 2, 3), (1, 2, 3))).astype(x.dtype, copy=False)
    if b is not None:
        y += b
    return Variable(numpy.rollaxis(y, 3, 1))


def relu(x):
    return Variable(numpy.maximum(_a(x), 0, dtype=_a(x).dtype))


def tanh(x):
    return Variable(numpy.tanh(_a(x)))


def sigmoid(x):
    x = _a(x)
    half = x.dtype.type(0.5)
    return Variable(numpy.tanh(x * half) * half + half)      # Chainer's CPU formulation


def exp(x):
    return Variable(numpy.exp(_a(x)))


def log(x):
    return Variable(numpy.log(_a(x)))


def softplus(x, beta=1.0):
    x = _a(x)
    bx = beta * x
    return Variable((numpy.maximum(bx, 0) + numpy.log1p(numpy.exp(-numpy.fabs(bx)))) / beta)


def maximum(a, b):
    return Variable(numpy.maximum(_a(a), _a(b)))


def where(cond, a, b):
    return Variable(numpy.where(_a(cond), _a(a), _a(b)))


def broadcast_to(x, shape):
    return Variable(numpy.broadcast_to(_a(x), shape))


def mean(x):
    return Variable(_a(x).mean())


def split_axis(x, indices_or_sections, axis):
    return tuple(Variable(a) for a in numpy.split(_a(x), indices_or_sections, axis))


def concat(xs, axis=1):
    return Variable(numpy.concatenate([_a(x) for x in xs], axis=axis))


def dropout(x, ratio=.5):
    return x


def log_softmax(x, axis=1):
    x = _a(x)
    m = x.max(axis=axis, keepdims=True)
    y = x - m
    return Variable(y - numpy.log(numpy.exp(y).sum(axis=axis, keepdims=True)))


def softmax(x, axis=1):
    x = _a(x)
    y = numpy.exp(x - x.max(axis=axis, keepdims=True))
    return Variable(y / y.sum(axis=axis, keepdims=True))


def logsumexp(x, axis=None):
    x = _a(x)
    m = x.max(axis=axis, keepdims=True)
    return Variable(numpy.log(numpy.exp(x - m).sum(axis=axis)) + numpy.squeeze(m, axis=axis))


def softmax_cross_entropy(x, t):
    x, t = _a(x), _a(t)
    logp = _a(log_softmax(x))
    logp = numpy.rollaxis(logp, 1, logp.ndim).reshape(-1, x.shape[1])
    tt = t.ravel()
    picked = logp[numpy.arange(tt.size), tt]
    return Variable((-picked.sum() / max(tt.size, 1)).astype(x.dtype))


def resize_images(x, output_shape):
    x = _a(x)
    out_H, out_W = output_shape
    B, C, H, W = x.shape
    u_1d = numpy.linspace(0, W - 1, num=out_W)
    v_1d = numpy.linspace(0, H - 1, num=out_H)
    grid = numpy.meshgrid(u_1d, v_1d)
    u = grid[0].ravel()
    v = grid[1].ravel()
    u0 = numpy.floor(u).astype(numpy.int32)
    u0 = u0.clip(0, W - 2)
    u1 = u0 + 1
    v0 = numpy.floor(v).astype(numpy.int32)
    v0 = v0.clip(0, H - 2)
    v1 = v0 + 1
    w1 = (u1 - u) * (v1 - v)
    w2 = (u - u0) * (v1 - v)
    w3 = (u1 - u) * (v - v0)
    w4 = (u - u0) * (v - v0)
    w1, w2, w3, w4 = (w.astype(x.dtype) for w in (w1, w2, w3, w4))
    y = (w1[None, None, :] * x[:, :, v0, u0] + w2[None, None, :] * x[:, :, v0, u1] +
         w3[None, None, :] * x[:, :, v1, u0] + w4[None, None, :] * x[:, :, v1, u1])
    return Variable(y.reshape(B, C, out_H, out_W))
