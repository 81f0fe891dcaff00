"""Plugging ``cuda_compute`` into the *unmodified* NumS host layers.

The reference registers compute back-ends in ``application_manager.create``
(/root/reference/nums/core/application_manager.py:61-80) and its tests build applications with
``get_app(mode)`` (/root/reference/tests/conftest.py:51-72).  This module is the code a maintainer
would put behind a ``"cuda"`` entry of both: it finds an installed NumS, imports it, and returns the
reference's own ``ArrayApplication`` / ``FileSystem`` / ``BlockArray`` running over ``CudaSystem`` +
``cuda_compute``.  Nothing of the reference is copied; ``BlockArray``, ``ArrayApplication``,
``nums.numpy`` and ``nums.models.glms`` run as they are.

Where the reference comes from (first hit wins): ``$NUMS_REFERENCE_ROOT``; an importable ``nums``;
``/root/reference`` (build container); ``<repo>/baseline/_ref`` (the pip ``--target`` install made by
``scripts/install_reference.sh``, which is what travels to the GPU box).

The reference pins ``numpy<=1.20`` and ``ray<1.1`` (``setup.py:20-25``); this image has numpy 2.3 and
no ray / boto3, so ``install_compat()`` provides

* inert stub modules for ``ray`` (``systems.py:22``), ``boto3`` (``storage.py:24``) and
  ``numpy.compat`` (``filesystem.py:22``) -- only the serial / cuda systems are usable;
* the NumPy aliases removed since 1.20 that the reference still spells (``np.int``, ``np.float``,
  ``np.bool``, ``np.object``, ``np.product``, ``np.NINF`` ...; e.g. ``base.py:38``,
  ``blockarray.py:49``, ``nums/numpy/api.py:48-50``).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
INSTALLED_ROOT = os.path.join(REPO_ROOT, "baseline", "_ref")


def _is_reference(root):
    return bool(root) and os.path.isdir(os.path.join(root, "nums", "core"))


def reference_root():
    """Directory to put on ``sys.path`` for ``import nums``, or None."""
    for root in (os.environ.get("NUMS_REFERENCE_ROOT"), "/root/reference", INSTALLED_ROOT):
        if _is_reference(root):
            return root
    return None


def available():
    if reference_root() is not None:
        return True
    try:
        import importlib.util
        return importlib.util.find_spec("nums") is not None
    except Exception:
        return False


def install_compat():
    """NumPy-2 / no-ray compatibility shims (idempotent)."""
    import numpy as np

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

    try:
        import numpy.compat as compat
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

    if "ray" not in sys.modules:
        try:
            import ray  # noqa: F401
        except Exception:
            ray = types.ModuleType("ray")
            ray.__path__ = []

            def _no_ray(*_a, **_k):
                raise RuntimeError("ray is not installed; use NUMS_SYSTEM=serial or cuda")

            for fn in ("init", "put", "get", "remote", "nodes"):
                setattr(ray, fn, _no_ray)
            ray.shutdown = lambda *a, **k: None
            ray.is_initialized = lambda: False
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
        try:
            import boto3  # noqa: F401
        except Exception:
            boto3 = types.ModuleType("boto3")
            boto3.resource = lambda *a, **k: None
            boto3.client = lambda *a, **k: None
            sys.modules["boto3"] = boto3


_nums = None


def load_reference(system_name="serial"):
    """Import the reference package (once) and return the ``nums`` module."""
    global _nums
    if _nums is not None:
        return _nums
    root = reference_root()
    install_compat()
    if root is not None and root not in sys.path:
        sys.path.insert(0, root)
    os.environ.setdefault("NUMS_SYSTEM", system_name)
    try:
        import nums
    except ImportError as exc:
        raise RuntimeError("NumS is not installed: set NUMS_REFERENCE_ROOT or run scripts/install_reference.sh "
                           "(%s)" % exc) from exc
    from nums.core import settings
    if settings.system_name not in ("serial", "cuda"):
        settings.system_name = system_name
    _register_cuda_system()
    if os.environ.get("NUMS_HOST_FAST_PATH", "1") != "0":
        _install_host_fast_path()
    _nums = nums
    return nums


def _install_host_fast_path():
    """SURVEY.md section 8f.2, the part that needs no source edit: the host layers infer every result dtype by
    running the ufunc on freshly built 0-d arrays (``nums/core/array/utils.py:27-52``) -- once per block per
    operation, ~8 us each, which is more than a quarter of the per-block dispatch budget when the kernel itself
    takes ~40 us (config 1).  The three helpers are pure functions of (op name, dtype[s]); they are replaced by
    memoised versions of themselves.  Results are identical by construction (the original function computes
    every cache entry)."""
    import functools
    from nums.core.array import utils as array_utils
    for name in ("get_bop_output_type", "get_uop_output_type", "get_reduce_output_type"):
        fn = getattr(array_utils, name)
        if getattr(fn, "_nums_b200", False):
            continue
        cached = functools.lru_cache(maxsize=None)(fn)

        def wrapper(*args, _cached=cached, _fn=fn):
            try:
                return _cached(*args)
            except TypeError:            # unhashable argument: fall through to the original
                return _fn(*args)
        wrapper._nums_b200 = True
        wrapper.__name__ = name
        setattr(array_utils, name, wrapper)


def _register_cuda_system():
    """The registry edit of INTEGRATION.md section 1, applied at run time instead of to the source:
    ``application_manager.create`` (application_manager.py:51-82) learns ``system_name == "cuda"``
    (``NUMS_SYSTEM=cuda``), so ``nums.init()`` and every lazily initialised ``nums.numpy`` call land on
    the GPU application."""
    from nums.core import application_manager as am
    if getattr(am.create, "_nums_b200", False):
        return
    reference_create = am.create

    def create():
        if am.settings.system_name != "cuda":
            return reference_create()
        if am._instance is not None:
            raise Exception("create() called more than once.")
        return cuda_app()
    create._nums_b200 = True
    am.create = create


_system_cls = None


def cuda_system_class():
    """``CudaSystem`` as a subclass of the reference's ``SerialSystem``.

    The host layers test ``isinstance(app.system, SerialSystem)`` in places
    (nums/numpy/numpy_utils.py:66-73, application.py:49-57); ``CudaSystem`` is an in-process system just
    like ``SerialSystem`` (same ``call`` / ``register`` / ``get_options`` contract, systems.py:68-142),
    only with ``put`` / ``get`` crossing PCIe, so it takes that place in the hierarchy.  All behaviour
    comes from ``CudaSystem`` (first in the MRO)."""
    global _system_cls
    if _system_cls is None:
        load_reference()
        from nums.core.systems.systems import SerialSystem
        from nums_b200.cuda_system import CudaSystem

        class ReferenceCudaSystem(CudaSystem, SerialSystem):
            # The reference's System routes kernel names through a Python-level __getattribute__ (systems.py:62-66),
            # which then intercepts EVERY attribute access of the system object (~1 us each, ~16 per kernel call
            # inside our dispatch code).  CudaSystem resolves kernel names in __getattr__ instead -- reached only
            # when normal lookup fails -- so plain lookups can take the C fast path again; no kernel name collides
            # with an attribute of the class (tests/test_host_logic.py).
            __getattribute__ = object.__getattribute__

        _system_cls = ReferenceCudaSystem
    return _system_cls


_spmd_cls = None


def spmd_system_class():
    """``SpmdSystem`` (nums_b200.spmd) in the same place of the reference's class hierarchy."""
    global _spmd_cls
    if _spmd_cls is None:
        load_reference()
        from nums.core.systems.systems import SerialSystem
        from nums_b200.spmd import SpmdSystem

        class ReferenceSpmdSystem(SpmdSystem, SerialSystem):
            __getattribute__ = object.__getattribute__      # see ReferenceCudaSystem

        _spmd_cls = ReferenceSpmdSystem
    return _spmd_cls


def cuda_system(**system_kwargs):
    """The system object for this process: ``CudaSystem`` on one GPU; under ``torchrun`` (WORLD_SIZE > 1, one
    process per GPU) an ``SpmdSystem`` that partitions the block grid over the ranks."""
    from nums_b200 import multi_gpu
    rank, world = multi_gpu.init_distributed()
    if world > 1:
        from nums_b200.cuda_system import CudaSystem
        local = CudaSystem(rank=rank, world_size=world, **system_kwargs)
        local.init()
        system = spmd_system_class()(local)
    else:
        system = cuda_system_class()(**system_kwargs)
    system.init()
    return system


def cuda_app(**system_kwargs):
    """The reference's ``ArrayApplication`` on this process' GPU(s) -- ``get_app("cuda")``."""
    load_reference()
    from nums.core.array.application import ArrayApplication
    from nums.core.systems.filesystem import FileSystem
    system = cuda_system(**system_kwargs)
    return ArrayApplication(system=system, filesystem=FileSystem(system))


def use_cuda(**system_kwargs):
    """Make the ``nums`` / ``nums.numpy`` module-level API run on the GPU
    (application_manager.py:38-48: ``set_instance``).  Returns the application."""
    load_reference()
    from nums.core import application_manager, settings
    if application_manager.is_initialized():
        application_manager.destroy()
    settings.system_name = "cuda"
    app = cuda_app(**system_kwargs)
    application_manager.set_instance(app)
    return app
