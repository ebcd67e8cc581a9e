from .variable import Variable, _unwrap


class _InTypes(tuple):
    def size(self):
        return len(self)


class FunctionNode:
    """apply() = unwrap -> forward(raw arrays) -> wrap.  backward() is invoked
    explicitly by the test harness (no graph traversal in the stub)."""

    def __init__(self):
        self._in_data = None
        self._out_data = None
        self._retain_in = ()
        self._retain_out = ()
        self.inputs = None

    def check_type_forward(self, in_types):
        pass

    def retain_inputs(self, indexes):
        self._retain_in = tuple(indexes)

    def retain_outputs(self, indexes):
        self._retain_out = tuple(indexes)

    def apply(self, inputs):
        self._in_data = tuple(_unwrap(x) for x in inputs)
        outs = self.forward(self._in_data)
        if not isinstance(outs, tuple):
            outs = (outs,)
        self._out_data = tuple(_unwrap(o) for o in outs)
        return tuple(Variable(o) for o in self._out_data)

    def get_retained_inputs(self):
        return tuple(Variable(self._in_data[i]) for i in self._retain_in)

    def get_retained_outputs(self):
        return tuple(Variable(self._out_data[i]) for i in self._retain_out)

    def forward(self, inputs):
        raise NotImplementedError

    def backward(self, target_input_indexes, grad_outputs):
        raise NotImplementedError
