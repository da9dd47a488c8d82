"""Import shim (test infrastructure): findiff.FinDiff(...).stencil(shape).data for acc=2, as used
at reference grad_utils.py:154-159 (init time only).  Closed-form second-order tables:
  d/dx    C: {-1:-1/2h, +1:+1/2h}      L: {0:-3/2h, 1:2/h, 2:-1/2h}     H: {0:3/2h, -1:-2/h, -2:1/2h}
  d2/dx2  C: {-1:1, 0:-2, 1:1}/h^2     L: {0:2, 1:-5, 2:4, 3:-1}/h^2    H: {0:2, -1:-5, -2:4, -3:-1}/h^2
Mixed derivatives are tensor products.  findiff itself is not installed here: parity of these
tables against findiff>=0.10 is pinned analytically only (exactness on quadratics), see DESIGN.md."""
import itertools


def _tab1d(order, h):
    if order == 1:
        return {'C': {-1: -0.5 / h, 1: 0.5 / h},
                'L': {0: -1.5 / h, 1: 2.0 / h, 2: -0.5 / h},
                'H': {0: 1.5 / h, -1: -2.0 / h, -2: 0.5 / h}}
    if order == 2:
        h2 = h * h
        return {'C': {-1: 1.0 / h2, 0: -2.0 / h2, 1: 1.0 / h2},
                'L': {0: 2.0 / h2, 1: -5.0 / h2, 2: 4.0 / h2, 3: -1.0 / h2},
                'H': {0: 2.0 / h2, -1: -5.0 / h2, -2: 4.0 / h2, -3: -1.0 / h2}}
    raise NotImplementedError(order)


class _StencilSet:
    def __init__(self, data):
        self.data = data


class FinDiff:
    def __init__(self, *args, acc=2):
        if acc != 2:
            raise NotImplementedError('shim supports acc=2 only (model.yaml: fd_acc: 2)')
        if isinstance(args[0], (tuple, list)):
            self.terms = [tuple(a) for a in args]
        else:
            self.terms = [tuple(args)]

    def stencil(self, shape):
        ndim = len(shape)
        data = {}
        for key in itertools.product('LCH', repeat=ndim):
            # product over the partial derivatives (one per listed axis)
            st = {tuple([0] * ndim): 1.0}
            for (axis, h, order) in self.terms:
                tab = _tab1d(order, h)[key[axis]]
                new = {}
                for off, c in st.items():
                    for o, c1 in tab.items():
                        off2 = list(off)
                        off2[axis] += o
                        new[tuple(off2)] = new.get(tuple(off2), 0.0) + c * c1
                st = new
            data[key] = st
        return _StencilSet(data)
