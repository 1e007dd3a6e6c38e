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


# ----------------------------------------------------------------------------------------
# Python lambda strings -> CUDA.  In YAML input an external likelihood / prior is a string
# such as ``"lambda a, b: stats.norm.logpdf(a - b**2, loc=0, scale=0.1) - np.log1p(a*a)"``
# (cobaya/tools.py:344-384 evaluates it with ``np`` and ``stats`` in scope).  Arithmetic
# expressions of that kind are translated to their CUDA twin automatically; anything outside
# the small grammar below raises ``DeviceFunctionError`` (the model is then refused, never
# evaluated on the CPU).  The plugin checks every translated function against the Python
# callable at the start points before sampling.
# ----------------------------------------------------------------------------------------
import ast
import math

_UNARY = {"exp": "exp", "log": "log", "log10": "log10", "log2": "log2", "log1p": "log1p",
          "expm1": "expm1", "sqrt": "sqrt", "sin": "sin", "cos": "cos", "tan": "tan",
          "arcsin": "asin", "arccos": "acos", "arctan": "atan", "asin": "asin", "acos": "acos",
          "atan": "atan", "sinh": "sinh", "cosh": "cosh", "tanh": "tanh", "abs": "fabs",
          "fabs": "fabs", "absolute": "fabs", "floor": "floor", "ceil": "ceil",
          "square": None, "erf": "erf", "erfc": "erfc", "lgamma": "lgamma", "cbrt": "cbrt"}
_BINARY = {"power": "pow", "pow": "pow", "arctan2": "atan2", "atan2": "atan2",
           "minimum": "fmin", "maximum": "fmax", "fmin": "fmin", "fmax": "fmax",
           "hypot": "hypot", "fmod": "fmod"}
_CONST = {"pi": math.pi, "e": math.e, "inf": math.inf}
_MODULES = {"np", "numpy", "math"}


class _Lambda2Cuda(ast.NodeVisitor):
    def __init__(self, args):
        self.args = {a: i for i, a in enumerate(args)}
        self.in_test = 0   # inside the condition of a conditional expression

    def bad(self, node, why):
        raise DeviceFunctionError(f"cannot translate to CUDA: {why} "
                                  f"({ast.unparse(node) if hasattr(ast, 'unparse') else node})")

    def visit(self, node):
        m = getattr(self, "v_" + type(node).__name__, None)
        if m is None:
            self.bad(node, f"unsupported syntax {type(node).__name__}")
        return m(node)

    def v_Constant(self, n):
        if isinstance(n.value, bool) or not isinstance(n.value, (int, float)):
            self.bad(n, "only numeric constants")
        v = float(n.value)
        if math.isinf(v):
            return "(-CB2_INF)" if v < 0 else "CB2_INF"
        return f"({v!r})" if v < 0 else repr(v)

    def v_Name(self, n):
        if n.id in self.args:
            return f"p[{self.args[n.id]}]"
        if n.id in ("abs", "min", "max", "pow", "float"):
            self.bad(n, "a function name used as a value")
        self.bad(n, f"unknown name '{n.id}' (only the lambda's own arguments)")

    def v_UnaryOp(self, n):
        x = self.visit(n.operand)
        if isinstance(n.op, ast.USub):
            return f"(-{x})"
        if isinstance(n.op, ast.UAdd):
            return x
        self.bad(n, "unary operator")

    def v_BinOp(self, n):
        a, b = self.visit(n.left), self.visit(n.right)
        if isinstance(n.op, ast.Add):
            return f"({a} + {b})"
        if isinstance(n.op, ast.Sub):
            return f"({a} - {b})"
        if isinstance(n.op, ast.Mult):
            return f"({a} * {b})"
        if isinstance(n.op, ast.Div):
            return f"({a} / {b})"
        if isinstance(n.op, ast.Pow):
            if isinstance(n.right, ast.Constant) and n.right.value in (2, 2.0):
                return f"cb2_sq({a})"
            return f"pow({a}, {b})"
        self.bad(n, "binary operator")

    def v_Compare(self, n):
        if len(n.ops) != 1:
            self.bad(n, "chained comparison")
        ops = {ast.Lt: "<", ast.LtE: "<=", ast.Gt: ">", ast.GtE: ">=", ast.Eq: "==",
               ast.NotEq: "!="}
        op = ops.get(type(n.ops[0]))
        if op is None:
            self.bad(n, "comparison operator")
        expr = f"({self.visit(n.left)} {op} {self.visit(n.comparators[0])})"
        return expr if self.in_test else f"({expr} ? 1.0 : 0.0)"

    def v_BoolOp(self, n):
        # Python's and/or return one of their operands; only as a truth value (the condition
        # of `x if cond else y`) do they mean what && / || mean
        if not self.in_test:
            self.bad(n, "and/or outside the condition of a conditional expression")
        op = " && " if isinstance(n.op, ast.And) else " || "
        return "(" + op.join(self.visit(v) for v in n.values) + ")"

    def v_IfExp(self, n):
        self.in_test += 1
        test = self.visit(n.test)
        self.in_test -= 1
        if not isinstance(n.test, (ast.Compare, ast.BoolOp)):
            test = f"({test} != 0.0)"
        return f"({test} ? {self.visit(n.body)} : {self.visit(n.orelse)})"

    def v_Attribute(self, n):   # np.pi, math.e, np.inf
        if isinstance(n.value, ast.Name) and n.value.id in _MODULES and n.attr in _CONST:
            v = _CONST[n.attr]
            return "CB2_INF" if math.isinf(v) else repr(v)
        self.bad(n, "attribute")

    def _fname(self, f):
        if isinstance(f, ast.Name):
            return None, f.id
        if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name):
            return f.value.id, f.attr
        if (isinstance(f, ast.Attribute) and isinstance(f.value, ast.Attribute)
                and isinstance(f.value.value, ast.Name)):
            return f"{f.value.value.id}.{f.value.attr}", f.attr
        return "?", "?"

    def v_Call(self, n):
        mod, name = self._fname(n.func)
        args = [self.visit(a) for a in n.args]
        kw = {k.arg: self.visit(k.value) for k in n.keywords}
        if mod == "stats.norm" and name == "logpdf":
            # scipy.stats.norm.logpdf(x, loc=0, scale=1)
            x = args[0] if args else self.bad(n, "missing argument")
            loc = args[1] if len(args) > 1 else kw.get("loc", "0.0")
            sc = args[2] if len(args) > 2 else kw.get("scale", "1.0")
            return (f"(-0.5 * cb2_sq(({x} - {loc}) / {sc}) - log({sc}) - "
                    f"{0.5 * math.log(2 * math.pi)!r})")
        if mod == "stats.uniform" and name == "logpdf":
            x = args[0] if args else self.bad(n, "missing argument")
            loc = args[1] if len(args) > 1 else kw.get("loc", "0.0")
            sc = args[2] if len(args) > 2 else kw.get("scale", "1.0")
            return (f"((({x}) >= ({loc}) && ({x}) <= ({loc}) + ({sc})) ? -log({sc}) : -CB2_INF)")
        if kw:
            self.bad(n, "keyword arguments")
        if mod is not None and mod not in _MODULES:
            self.bad(n, f"function of module '{mod}'")
        if mod is None and name in ("abs",) and len(args) == 1:
            return f"fabs({args[0]})"
        if mod is None and name in ("min", "max") and len(args) == 2:
            return f"{'fmin' if name == 'min' else 'fmax'}({args[0]}, {args[1]})"
        if mod is None and name == "pow" and len(args) == 2:
            return f"pow({args[0]}, {args[1]})"
        if mod is None and name == "float" and len(args) == 1:
            return args[0]
        if mod is None:
            self.bad(n, f"call of '{name}'")
        if name in _UNARY and len(args) == 1:
            return f"cb2_sq({args[0]})" if name == "square" else f"{_UNARY[name]}({args[0]})"
        if name in _BINARY and len(args) == 2:
            return f"{_BINARY[name]}({args[0]}, {args[1]})"
        self.bad(n, f"function '{mod}.{name}' with {len(args)} argument(s)")


def cuda_from_lambda(source: str, name: str):
    """Translate ``"lambda a, b: <arithmetic expression>"`` into CUDA.  Returns
    ``(cuda_source, argument names)``; raises ``DeviceFunctionError`` outside the grammar."""
    if not isinstance(source, str):
        raise DeviceFunctionError("not a lambda string")
    try:
        tree = ast.parse(source.strip(), mode="eval").body
    except SyntaxError as e:
        raise DeviceFunctionError(f"cannot parse {source!r}: {e}") from e
    if not isinstance(tree, ast.Lambda):
        raise DeviceFunctionError("only `lambda ...: expression` strings are translated")
    a = tree.args
    if a.vararg or a.kwarg or a.kwonlyargs or a.defaults or getattr(a, "posonlyargs", None):
        raise DeviceFunctionError("lambda with defaults / *args / **kwargs")
    args = [x.arg for x in a.args]
    if "_self" in args or not args:
        raise DeviceFunctionError("lambda without parameters, or using `_self`")
    ident = re.sub(r"\W", "_", name)
    if not re.match(r"[A-Za-z_]", ident):
        ident = "f_" + ident
    body = _Lambda2Cuda(args).visit(tree.body)
    src = ("#define CB2_INF (__longlong_as_double(0x7ff0000000000000LL))\n"
           "__device__ __forceinline__ double cb2_sq(double x) { return x * x; }\n"
           f"// translated from: {source.strip()}\n"
           f'extern "C" __device__ double cb2_fn_{ident}(const double *p, int n) {{\n'
           f"    return {body};\n}}\n")
    return src, f"cb2_fn_{ident}", args


def twin_of(function, source_text, name):
    """(cuda_source, entry point) of an external function: the attached twin
    (``device_function``), else the translation of its lambda string; else an error."""
    if getattr(function, "cuda_source", None):
        return function.cuda_source, getattr(function, "cuda_name", None) or function.__name__
    if isinstance(source_text, str):
        src, entry, _ = cuda_from_lambda(source_text, name)
        return src, entry
    raise DeviceFunctionError(
        "a Python callable without a CUDA twin (cobaya_b200.functor.device_function) that is "
        "not given as a translatable lambda string")
