"""Import shim (test infrastructure): matplotlib is not installed in this image.  The reference's modules import
pyplot at import time (denoising_utils.py:8-29) and its driver scripts plot samples between training iterations
(main.py:240-262, main_toy.py:175-199).  Every plotting call is accepted and ignored: figures are not part of any
parity or performance claim."""


class _Anything:
    """object that accepts any attribute access / call / indexing / iteration-free use and returns itself"""

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __getitem__(self, i):
        return self

    def __iter__(self):
        return iter(())


class _Style:
    def use(self, *a, **k):
        pass


style = _Style()


def rc(*a, **k):
    pass


def subplots(nrows=1, ncols=1, *a, **k):
    fig = _Anything()
    if nrows * ncols == 1:
        return fig, _Anything()
    return fig, [_Anything() for _ in range(nrows * ncols)]


def __getattr__(name):
    return _Anything()
