"""
Device functors: the CUDA side of an external likelihood function.

The reference takes any Python callable as a likelihood (``LikelihoodExternalFunction``,
cobaya/likelihood.py:150-255).  A Python callable cannot run inside a GPU kernel; the engine
takes its CUDA twin instead, attached to the callable so that the *same* ``info`` runs on the
reference (Python function) and on the engine (CUDA function)::

    from cobaya_b200.functor import device_function

    @device_function('''
    extern "C" __device__ double banana(const double *p, int n) {
        const double a = p[0], b = p[1];
        return -0.5 * (a * a + 10.0 * (b - a * a) * (b - a * a));
    }''')
    def banana(a, b):
        return -0.5 * (a**2 + 10 * (b - a**2) ** 2)

    info = {"likelihood": {"banana": banana}, "params": {"a": ..., "b": ...},
            "sampler": {"cobaya_b200.plugin.MCMC": None}}

``p`` holds the function's arguments in the order of the Python signature
(= ``Likelihood.input_params``); the function returns log L.  The source is compiled with
NVRTC when the engine is built (``cb2_add_external_likelihood``); the plugin then checks the
two implementations against each other at the start points and refuses a mismatch.
"""

from __future__ import annotations

import ctypes as C
import re


class DeviceFunctionError(ValueError):
    pass


def _entry_point(source: str, name: str | None):
    if name:
        return name
    m = re.findall(r'extern\s+"C"\s+__device__\s+double\s+([A-Za-z_]\w*)\s*\(', source)
    if len(m) != 1:
        raise DeviceFunctionError(
            'the CUDA source must define exactly one `extern "C" __device__ double NAME('
            "const double *p, int n)` (or pass name=...)")
    return m[0]


def device_function(cuda_source: str, name: str | None = None):
    """Decorator attaching ``cuda_source`` / ``cuda_name`` to a Python likelihood function."""
    entry = _entry_point(cuda_source, name)

    def deco(fn):
        fn.cuda_source = cuda_source
        fn.cuda_name = entry
        return fn

    return deco


def check_source(cuda_source: str, name: str | None = None, dim: int = 1):
    """Compile ``cuda_source`` with NVRTC (no GPU needed).  Returns the compiler log (warnings);
    raises ``DeviceFunctionError`` with the log on errors."""
    from . import _cabi

    lib = _cabi.load()
    entry = _entry_point(cuda_source, name)
    log = C.create_string_buffer(1 << 14)
    rc = lib.cb2_check_external_source(cuda_source.encode(), entry.encode(), int(dim), log,
                                       len(log))
    text = log.value.decode(errors="replace")
    if rc != 0:
        raise DeviceFunctionError(f"external function '{entry}' did not compile ({rc}):\n{text}")
    return text
