import contextlib

import numpy

from .variable import Parameter


class Link(object):
    def __init__(self):
        self._params = []
        self._within = False
        self.name = None

    @property
    def xp(self):
        return numpy

    @contextlib.contextmanager
    def init_scope(self):
        old = self._within
        self._within = True
        try:
            yield
        finally:
            self._within = old

    def __setattr__(self, name, value):
        if getattr(self, "_within", False) and isinstance(value, Parameter):
            value.name = name
            self._params.append(name)
        super(Link, self).__setattr__(name, value)

    def namedparams(self, include_uninit=True):
        for name in sorted(self._params):
            yield "/" + name, getattr(self, name)

    def params(self, include_uninit=True):
        for _, p in self.namedparams():
            yield p

    def namedlinks(self):
        yield "/", self

    def children(self):
        return iter(())

    def cleargrads(self):
        pass

    def __call__(self, *a, **k):
        raise NotImplementedError


class Chain(Link):
    def __init__(self):
        super(Chain, self).__init__()
        self._children = []

    def __setattr__(self, name, value):
        if getattr(self, "_within", False) and isinstance(value, Link):
            value.name = name
            self._children.append(name)
        super(Chain, self).__setattr__(name, value)

    def children(self):
        for name in sorted(self._children):
            yield getattr(self, name)

    def namedparams(self, include_uninit=True):
        for ret in super(Chain, self).namedparams():
            yield ret
        for name in sorted(self._children):
            prefix = "/" + name
            for path, p in getattr(self, name).namedparams():
                yield prefix + path, p


class ChainList(Link):
    def __init__(self, *links):
        super(ChainList, self).__init__()
        self._list = []
        for l in links:
            self.add_link(l)

    def add_link(self, link):
        link.name = str(len(self._list))
        self._list.append(link)

    def children(self):
        return iter(self._list)

    def __iter__(self):
        return iter(self._list)

    def __len__(self):
        return len(self._list)

    def __getitem__(self, i):
        return self._list[i]

    def namedparams(self, include_uninit=True):
        for ret in super(ChainList, self).namedparams():
            yield ret
        for i, l in enumerate(self._list):
            for path, p in l.namedparams():
                yield "/%d%s" % (i, path), p
