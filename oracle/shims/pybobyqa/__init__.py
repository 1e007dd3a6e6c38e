"""Import shim (TEST INFRASTRUCTURE ONLY): the reference's ``minimize`` sampler imports
``pybobyqa`` at module level (cobaya/samplers/minimize/minimize.py:105-106); the package is not
in this image and there is no network.  The tests use ``method: scipy``; asking for BOBYQA
through this shim fails loudly."""
from . import controller  # noqa: F401


def solve(*args, **kwargs):
    raise ImportError("pybobyqa is not installed (import shim of the test-suite)")
