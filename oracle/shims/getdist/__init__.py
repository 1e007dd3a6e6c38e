"""Minimal stand-in for the `getdist` package (TEST/ORACLE INFRASTRUCTURE ONLY).

The reference (cobaya/collection.py:18-19) imports `getdist` at module import
time; getdist is not installed in this image and there is no network.  This
shim provides just the names needed to *import* the reference so its own
functions can be driven as the parity oracle.  `MCSamples.confidence`
(weighted quantiles, used by the R-1-of-bounds stopping rule) is NOT provided:
runs that reach that branch raise.
"""


class MCSamples:  # pragma: no cover - placeholder
    def __init__(self, *a, **k):
        raise NotImplementedError("getdist shim: MCSamples is a placeholder")


from . import chains  # noqa: E402,F401
