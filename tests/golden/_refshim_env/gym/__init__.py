"""Import-only stand-in for `gym` (absent from this image): lets the reference's endtoend.py be
imported so that its numeric methods can be called.  No behaviour of the reference depends on it."""
from . import utils  # noqa: F401


class Env(object):
    pass


class _Spaces(object):
    class Box(object):
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Dict(dict):
        pass


spaces = _Spaces()
