"""Import shim (test infrastructure): the reference imports matplotlib at module import time only
(denoising_utils.py:8-29); no plotting happens on the measured path."""


class _RC(dict):
    def update(self, *a, **k):
        pass


rcParams = _RC()
