import numpy as _np


def get_array_module(*args):
    return _np
