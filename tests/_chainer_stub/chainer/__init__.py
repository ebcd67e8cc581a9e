"""Forward-only stand-in for the ~30 Chainer symbols the LQR/MPC hot path touches.

TEST INFRASTRUCTURE ONLY.  Chainer 6.3.0 cannot be installed in the build image
(no network), so this stub lets (a) the *unmodified* reference modules under
/root/reference be imported as a live oracle when generating golden vectors
(tests/golden/make_golden.py) and (b) the facade classes of this repo be
exercised through the FunctionNode.apply protocol.  It implements no autograd:
`FunctionNode.backward` is called explicitly by the harness.
"""
import contextlib

import numpy as _np

from . import backend, functions, function_node, utils  # noqa: F401
from .variable import Variable, Parameter, as_variable  # noqa: F401
from .link import Link, Chain  # noqa: F401

__version__ = "0.0-stub"
__is_stub__ = True


@contextlib.contextmanager
def no_backprop_mode():
    yield


@contextlib.contextmanager
def using_config(name, value):
    yield


def grad(*args, **kwargs):
    raise NotImplementedError("chainer stub is forward-only (no autograd)")
