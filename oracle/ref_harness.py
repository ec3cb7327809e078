"""Harness that imports and drives the UNMODIFIED reference (peckhams/topoflow36)
in this container.  TEST INFRASTRUCTURE ONLY -- nothing in the product package
imports this module, and on the GPU box it only ever reads the staged copy
under oracle/_ref/ (``make -C oracle _ref``; /root/reference does not exist there).

It is used (a) by ``oracle/gen_golden.py`` to produce the committed fixtures
under ``tests/golden/`` and (b) by the ``-m "not gpu"`` tests, when
``/root/reference`` is present, to pin the oracle restatements
(``oracle/np_oracle.py``, ``oracle/tf_oracle.c``) against live reference output.

What it has to work around; none of it touches
arithmetic:
  * optional imports that are absent here (netCDF4, osgeo, matplotlib, imageio)
    are stubbed with inert modules;
  * the reference tree is read-only and its examples write next to themselves,
    so examples are staged into a writable scratch copy;
  * ``ice_base`` (disabled path) forgets ``h_ice`` which ``met_base`` now wants.
"""
import os
import shutil
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """The live read-only tree in the build container; on the GPU box (where it does
    not exist) the archive that ``make -C oracle _ref`` staged as oracle/_ref/topoflow_ref.tar.gz,
    unpacked into a scratch directory for the life of the process."""
    env = os.environ.get("TF_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/topoflow/components"):
        return "/root/reference"
    tar = os.path.join(_HERE, "_ref", "topoflow_ref.tar.gz")
    if os.path.isfile(tar):
        import atexit
        import tarfile
        import tempfile
        tmp = tempfile.mkdtemp(prefix="tf_ref_")
        atexit.register(shutil.rmtree, tmp, ignore_errors=True)
        with tarfile.open(tar) as t:
            t.extractall(tmp, filter='data')
        return tmp
    return "/root/reference"


REF_ROOT = _find_root()


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "topoflow", "components"))


class _Inert(object):
    """Swallows every call/attribute/item access; stands in for netCDF4 etc."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __setattr__(self, name, value):
        pass

    def __setitem__(self, k, v):
        pass

    def __getitem__(self, k):
        return _Inert()

    def __len__(self):
        return 0

    def __iter__(self):
        return iter(())

    def __add__(self, o):
        return self

    __radd__ = __iadd__ = __sub__ = __mul__ = __add__

    def __int__(self):
        return 0

    __index__ = __int__

    def __float__(self):
        return 0.0


def _inert_module(name):
    m = types.ModuleType(name)
    m.__path__ = []
    m.__getattr__ = lambda attr: _Inert()   # module-level __getattr__ (PEP 562)
    return m


def install_import_shims():
    """Register inert stand-ins for the optional, off-hot-path dependencies."""
    try:
        # torch must finish its own first import BEFORE the stand-ins exist: its lazy
        # sub-imports probe optional packages (matplotlib among them) and choke on an
        # inert module
        import torch  # noqa: F401
    except ImportError:
        pass
    if "netCDF4" not in sys.modules:
        try:
            import netCDF4  # noqa: F401
        except Exception:
            m = _inert_module("netCDF4")
            m.__version__ = "0.0-shim"
            m.Dataset = _Inert
            m.date2num = lambda *a, **k: 0.0
            m.num2date = lambda *a, **k: None
            sys.modules["netCDF4"] = m
    if "osgeo" not in sys.modules:
        try:
            import osgeo  # noqa: F401
        except Exception:
            m = _inert_module("osgeo")
            for sub in ("gdal", "osr", "ogr", "gdalconst"):
                setattr(m, sub, _Inert())
                sys.modules["osgeo." + sub] = _inert_module("osgeo." + sub)
            sys.modules["osgeo"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.dates",
                 "matplotlib.colors", "matplotlib.cm", "matplotlib.ticker",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _inert_module(name)


_IMPORTED = False


def import_reference(quiet=True):
    """Put the reference on sys.path and import the ``topoflow`` package."""
    global _IMPORTED
    if not reference_available():
        raise RuntimeError("reference tree not found at " + REF_ROOT)
    install_import_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if not _IMPORTED:
        if quiet:
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):
                import topoflow  # noqa: F401
        else:
            import topoflow  # noqa: F401
        _apply_external_bugfixes()
        _IMPORTED = True
    import topoflow
    return topoflow


def _apply_external_bugfixes():
    """Monkeypatches applied from OUTSIDE the reference (never edits files)."""
    from topoflow.components import ice_base
    _orig = ice_base.ice_component.initialize

    def _init(self, *a, **k):
        r = _orig(self, *a, **k)
        if not hasattr(self, "h_ice"):
            self.h_ice = self.initialize_scalar(0)
        return r

    ice_base.ice_component.initialize = _init


def stage_treynor(work_dir):
    """Writable copy of the Treynor example laid out the way the reference wants.

    Returns (cfg_dir, out_dir).  HOME is pointed at ``work_dir/home`` because
    the example's path_info.cfg hard-wires ``~/TF_Output/Treynor``.
    """
    src = os.path.join(REF_ROOT, "topoflow", "examples", "Treynor_Iowa_30m")
    dst = os.path.join(work_dir, "Treynor")
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst)
    for root, dirs, files in os.walk(dst):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    cfg_dir = os.path.join(dst, "__No_Infil_June_20_67_Rain") + os.sep
    met = os.path.join(dst, "__met")
    os.makedirs(met, exist_ok=True)
    shutil.copy(os.path.join(cfg_dir, "June_20_67_rain_rates.txt"), met)
    home = os.path.join(work_dir, "home")
    out_dir = os.path.join(home, "TF_Output", "Treynor")
    os.makedirs(out_dir, exist_ok=True)
    os.environ["HOME"] = home
    return cfg_dir, out_dir + os.sep


def quiet_call(fn, *a, **k):
    import contextlib
    import io
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r
