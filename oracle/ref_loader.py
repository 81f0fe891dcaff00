"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (merrymercy/nums).

This module is part of ``oracle/``: only ``tests/``, ``oracle/make_golden.py``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import
it.  Nothing under ``nums_b200/`` may.

It imports the unmodified reference package from ``/root/reference`` (which
exists only in the build container, never on the GPU box) under a small
in-memory compatibility layer, because the reference pins ``numpy<=1.20`` and
``ray<1.1`` (``setup.py:20-25``) while this image ships numpy 2.3 and no ray /
boto3:

* stub modules for ``ray`` (``nums/core/systems/systems.py:22``),
  ``boto3`` (``nums/core/storage/storage.py:24``) and ``numpy.compat``
  (``nums/core/systems/filesystem.py:22``);
* the removed NumPy aliases the reference still uses (``np.int``, ``np.float``,
  ``np.bool``, ``np.object``, ``np.product``, ``np.NINF`` ... e.g.
  ``nums/core/array/base.py:38``, ``blockarray.py:49``, ``nums/numpy/api.py:48-50``).

Nothing of the reference is copied: the loader only puts ``/root/reference`` on
``sys.path``.  ``available()`` tells callers whether the reference is present.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NUMS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nums", "core"))


def _install_stubs():
    import numpy as np

    # -- removed numpy aliases -------------------------------------------------
    for name, val in (("int", int), ("float", float), ("bool", np.bool_),
                      ("object", object), ("complex", complex), ("str", str)):
        if name not in np.__dict__:
            setattr(np, name, val)
    if "product" not in np.__dict__:
        np.product = np.prod
    for name, val in (("NINF", -np.inf), ("PINF", np.inf), ("PZERO", 0.0),
                      ("NZERO", -0.0), ("Inf", np.inf), ("Infinity", np.inf),
                      ("NaN", np.nan), ("NAN", np.nan), ("infty", np.inf)):
        if name not in np.__dict__:
            setattr(np, name, val)
    st = np.lib.stride_tricks
    if not hasattr(st, "broadcast_to"):
        st.broadcast_to = np.broadcast_to

    # -- numpy.compat ------------------------------------------------------------
    try:
        import numpy.compat as compat  # still a namespace on some builds
    except Exception:  # pragma: no cover
        compat = types.ModuleType("numpy.compat")
        sys.modules["numpy.compat"] = compat
    if not hasattr(compat, "asbytes"):
        compat.asbytes = lambda s: s if isinstance(s, bytes) else str(s).encode("latin1")
        compat.asstr = lambda s: s.decode("latin1") if isinstance(s, bytes) else str(s)
        compat.asunicode = compat.asstr
        compat.os_fspath = os.fspath
        compat.contextlib_nullcontext = __import__("contextlib").nullcontext
        compat.is_pathlib_path = lambda p: hasattr(p, "__fspath__")
        sys.modules["numpy.compat"] = compat

    # -- ray / boto3 -------------------------------------------------------------
    if "ray" not in sys.modules:
        ray = types.ModuleType("ray")
        ray.__path__ = []

        def _no_ray(*_a, **_k):
            raise RuntimeError("ray is not installed; the oracle runs NUMS_SYSTEM=serial")

        for fn in ("init", "shutdown", "put", "get", "remote", "nodes", "is_initialized"):
            setattr(ray, fn, _no_ray)
        raylet = types.ModuleType("ray._raylet")
        raylet.ObjectRef = type("ObjectRef", (), {})
        ray._raylet = raylet
        actor = types.ModuleType("ray.actor")
        actor.ActorHandle = type("ActorHandle", (), {})
        ray.actor = actor
        ray.ObjectRef = raylet.ObjectRef
        sys.modules["ray"] = ray
        sys.modules["ray._raylet"] = raylet
        sys.modules["ray.actor"] = actor
    if "boto3" not in sys.modules:
        boto3 = types.ModuleType("boto3")
        boto3.resource = lambda *a, **k: None
        boto3.client = lambda *a, **k: None
        sys.modules["boto3"] = boto3


_loaded = None


def load():
    """Import the reference and return the ``nums`` package (serial system only)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    os.environ.setdefault("NUMS_SYSTEM", "serial")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import nums  # noqa: F401
    from nums.core import settings
    settings.system_name = "serial"
    _loaded = nums
    return nums


def serial_app(compute_module=None):
    """The reference's own ``ArrayApplication`` over ``SerialSystem``.

    With ``compute_module=None`` this is the reference oracle
    (``numpy_compute``); passing ``nums_b200.cuda_compute`` plugs the new
    backend into the unmodified reference host layers (tests/conftest.py:51-57
    of the reference builds its apps the same way).
    """
    load()
    from nums.core.systems import numpy_compute
    from nums.core.systems.systems import SerialSystem
    from nums.core.systems.filesystem import FileSystem
    from nums.core.array.application import ArrayApplication
    system = SerialSystem(compute_module=compute_module or numpy_compute)
    system.init()
    return ArrayApplication(system=system, filesystem=FileSystem(system))
