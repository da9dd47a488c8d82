class _Style:
    def use(self, *a, **k):
        pass


style = _Style()


def rc(*a, **k):
    pass


def __getattr__(name):
    def _f(*a, **k):
        raise RuntimeError('matplotlib shim: plotting is outside the hot path')
    return _f
