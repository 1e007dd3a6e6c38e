"""Optional access to the reference package (cobaya 3.6.2 installed under baseline/_ref by
`pip install --target`, plus the getdist import shim).  baseline/_ref is git-ignored but
travels to the GPU box; tests that need it skip when it is absent."""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REF, "cobaya"))


def enable_reference():
    if not have_reference():
        import pytest

        pytest.skip("reference package not installed under baseline/_ref")
    for p in (os.path.join(ROOT, "oracle", "shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging

    logging.getLogger().setLevel(logging.ERROR)
    import cobaya  # noqa: F401

    return cobaya
