"""CHECKER / BASELINE INFRASTRUCTURE (tests, bench.py --impl reference, tools/) — loaders for
the UNMODIFIED reference as installed by baseline/build_reference.sh (git-ignored
baseline/_ref, shipped to the GPU box).  Imports nothing of this repository's package.

* ``stock()``     the reference's `diff_gaussian_rasterization` package, built by its own
                  setup.py (its __init__.py + _C = ext.cpp + rasterize_points.cu + cuda_rasterizer/*)
* ``adapter()``   the reference's own __init__.py + rasterize_points.cu + ext.cpp, unmodified,
                  linked against libb3gs.so through csrc/reference_adapter.cpp
                  (baseline/build_adapter.py)
* ``render_adapter(dgr)``  the reference's `gaussian_renderer/__init__.py` (render()) and
                  `scene/gaussian_model.py` (GaussianModel) imported with `dgr` standing in for the
                  module name `diff_gaussian_rasterization` they import (gaussian_renderer/__init__.py:14)

Each loader returns None when its files are absent.  Modules are loaded under private
names so that the stock package, the adapter build and this repository's drop-in can live
in one process.
"""
import importlib
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
_cache = {}


def _stub_attr(attr):
    if attr.startswith("__"):          # inspect.getmodule() walks sys.modules asking for __file__ etc.
        raise AttributeError(attr)
    return object


def _load_package(alias, pkg_dir):
    init = os.path.join(pkg_dir, "__init__.py")
    if not os.path.exists(init) or not any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(pkg_dir)):
        return None
    if alias not in _cache:
        spec = importlib.util.spec_from_file_location(alias, init, submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        spec.loader.exec_module(mod)
        _cache[alias] = mod
    return _cache[alias]


def stock():
    return _load_package("dgr_stock_reference", os.path.join(REF_DIR, "diff_gaussian_rasterization"))


def adapter():
    return _load_package("dgr_reference_on_b3gs", os.path.join(REF_DIR, "adapter", "diff_gaussian_rasterization"))


def tree_root():
    for p in ("/root/reference", os.path.join(REF_DIR, "reference_tree")):
        if os.path.isfile(os.path.join(p, "gaussian_renderer", "__init__.py")):
            return p
    return None


def reference_module(name):
    """A module of the reference tree (e.g. "utils.loss_utils"), after render_adapter() has put
    the tree on sys.path and stubbed what the image lacks."""
    return importlib.import_module(name)


def render_adapter(dgr, alias):
    """The reference's gaussian_renderer module bound to the rasterizer package `dgr`.
    Returns (module with .render, GaussianModel class) or None."""
    root = tree_root()
    if root is None:
        return None
    key = ("render", alias)
    if key in _cache:
        return _cache[key]
    # scene/gaussian_model.py:18,20 imports plyfile and simple_knn._C (file IO and the k-NN
    # initialiser: neither is on the render path); stub what the image lacks
    # (never imported for real: this repository's top-level simple_knn/ shim would pull the product
    # package into a process that must stay free of it — the reference arm of bench.py)
    installed = []
    for name in ("plyfile", "simple_knn", "simple_knn._C"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.PlyData = m.PlyElement = object
            m.distCUDA2 = lambda *a, **k: None
            m.__path__ = []
            sys.modules[name] = m
            installed.append(name)
    if root not in sys.path:
        sys.path.insert(0, root)
    saved = sys.modules.get("diff_gaussian_rasterization")
    sys.modules["diff_gaussian_rasterization"] = dgr
    try:
        path = os.path.join(root, "gaussian_renderer", "__init__.py")
        for _ in range(30):
            try:
                spec = importlib.util.spec_from_file_location(alias, path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                break
            except ModuleNotFoundError as e:      # imageio, matplotlib, ...: pulled in by scene/__init__.py only
                assert not e.name.startswith(("diff_gaussian_rasterization", "binocular3dgs_b200", "gaussian_renderer"))
                stub = types.ModuleType(e.name)
                stub.__path__ = []
                stub.__getattr__ = _stub_attr
                sys.modules[e.name] = stub
                for k in [k for k in sys.modules if k.split(".")[0] in ("scene", "utils", "arguments")]:
                    del sys.modules[k]
        else:
            raise ImportError("could not import the reference's gaussian_renderer")
    finally:
        if saved is None:
            sys.modules.pop("diff_gaussian_rasterization", None)
        else:
            sys.modules["diff_gaussian_rasterization"] = saved
        for name in installed:      # the reference's modules have bound what they import; leave no stub behind
            sys.modules.pop(name, None)
    assert mod.GaussianRasterizer is dgr.GaussianRasterizer
    _cache[key] = (mod, mod.GaussianModel)
    return _cache[key]
