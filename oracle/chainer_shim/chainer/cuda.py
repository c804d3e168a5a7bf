import numpy


def get_array_module(*args):
    return numpy


def to_cpu(x):
    return x
