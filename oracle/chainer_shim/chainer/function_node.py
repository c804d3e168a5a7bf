from .variable import Variable, _a


class FunctionNode(object):
    """apply() = type check + forward on raw arrays (utils.py:161-236 relies on this)."""

    def retain_inputs(self, idx):
        pass

    def apply(self, inputs):
        arrays = tuple(_a(x) for x in inputs)
        if hasattr(self, "check_type_forward"):
            from .utils import type_check
            self.check_type_forward(type_check.TypeInfoTuple(arrays))
        outs = self.forward(arrays)
        return tuple(Variable(o) for o in outs)
