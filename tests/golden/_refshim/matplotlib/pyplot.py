"""No-op pyplot: every attribute is a function that does nothing."""


def __getattr__(name):
    def _noop(*args, **kwargs):
        return None
    return _noop
