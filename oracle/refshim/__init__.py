"""Run the reference's OWN, UNMODIFIED source files for the hot path in this container.

TEST INFRASTRUCTURE ONLY: imported by ``tests/golden/make_golden.py`` (fixture generation) and ``tests/test_reference_shim.py``; never by
the product, never on the GPU box (``/root/reference`` does not exist there -- the committed fixtures under ``tests/golden/`` travel instead).

The reference is pure Python on TensorFlow 1.x + GPflow 0.4.0, neither installable here.  ``install()`` registers two stand-in modules in
``sys.modules`` -- ``tensorflow`` (tf1.py: ~45 primitives mapped one-to-one onto torch.float64) and ``gpflow`` (gpflow04.py: the Param / Model
plumbing; its three arithmetic entry points are delegated to the reference's in-tree "Taken from GPFlow" copies) -- plus an inert ``matplotlib``
when that is absent, and then loads, by file path and under alias names (the product mirrors reuse the names ``onoffgpf`` / ``onofftf``):

    ref.main      <- /root/reference/onofftf/main.py        KernSE, GaussKL, GPConditional, GaussKLkron, tf_kron, Param, DataSet
    ref.onoffgpf  <- /root/reference/onoffgpf/              OnOffSVGP, OnOffLikelihood  (the GPflow model class itself)
    ref.utils     <- /root/reference/onofftf/utils.py       kernse_np

``script_functions`` compiles the nested model functions of a training script (scripts/onoff.py:143-280 etc.) from the script's own source
text at run time (``ast`` extraction of the ``def`` nodes; nothing is copied into this repository) so that they can be called eagerly.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ZIGP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "onofftf", "main.py"))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """register the stand-in ``tensorflow`` / ``gpflow`` (/ ``matplotlib``) modules; idempotent"""
    global _installed
    if _installed:
        return
    from . import gpflow04 as g, tf1
    if "tensorflow" in sys.modules and sys.modules["tensorflow"] is not tf1:
        raise RuntimeError("a real tensorflow is already imported; the reference shim must not shadow it")
    sys.modules["tensorflow"] = tf1
    _module("tensorflow.python")
    _module("tensorflow.python.framework")
    _module("tensorflow.python.framework.random_seed", get_seed=tf1.get_seed)
    gp = _module("gpflow", transforms=g.transforms, conditionals=g.conditionals, kullback_leiblers=g.kullback_leiblers, kernels=g.kernels)
    gp.param = _module("gpflow.param", Param=g.Param, AutoFlow=g.AutoFlow, DataHolder=g.DataHolder, Parameterized=g.Parameterized)
    gp.model = _module("gpflow.model", Model=g.Model)
    gp.mean_functions = _module("gpflow.mean_functions", Zero=g.Zero)
    gp._settings = _module("gpflow._settings", settings=g.settings)
    gp.minibatch = _module("gpflow.minibatch", MinibatchData=g.MinibatchData)
    gp.likelihoods = _module("gpflow.likelihoods", Likelihood=g.Likelihood)
    for sub, obj in (("transforms", g.transforms), ("conditionals", g.conditionals), ("kullback_leiblers", g.kullback_leiblers), ("kernels", g.kernels)):
        _module("gpflow." + sub, **vars(obj))
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        inert = lambda *a, **k: None
        mpl = _module("matplotlib", rcParams={}, use=inert)
        mpl.pyplot = _module("matplotlib.pyplot", switch_backend=inert)
        mpl.gridspec = _module("matplotlib.gridspec")
        mpl.ticker = _module("matplotlib.ticker")
    _installed = True


def _load(alias, path, package=False):
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, path, submodule_search_locations=[os.path.dirname(path)] if package else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


class _Ref:
    @property
    def main(self):
        install()
        return _load("zigp_ref_onofftf_main", os.path.join(REFERENCE_ROOT, "onofftf", "main.py"))

    @property
    def utils(self):
        return _load("zigp_ref_onofftf_utils", os.path.join(REFERENCE_ROOT, "onofftf", "utils.py"))

    @property
    def onoffgpf(self):
        install()
        return _load("zigp_ref_onoffgpf", os.path.join(REFERENCE_ROOT, "onoffgpf", "__init__.py"), package=True)


ref = _Ref()


def script_functions(relpath: str, outer: str, names, env: dict) -> dict:
    """Compile the nested functions ``names`` of function ``outer`` in the reference script ``relpath`` (source read from the reference tree at
    run time) into ``env`` -- the free variables those functions read (tf, np, float_type, jitter_level, num_inducing_f, ...) come from ``env``."""
    install()
    path = os.path.join(REFERENCE_ROOT, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == outer)
    picked = [n for n in fn.body if isinstance(n, ast.FunctionDef) and n.name in set(names)]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise KeyError("not found in %s:%s: %s" % (relpath, outer, sorted(missing)))
    code = compile(ast.Module(body=picked, type_ignores=[]), path, "exec")
    exec(code, env)
    return env
