"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (merrymercy/nums).

This module is part of ``oracle/``: only ``tests/``, ``oracle/make_golden.py``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg may import
it.  Nothing under ``nums_b200/`` may.

It imports the unmodified reference package from ``/root/reference`` (build
container) or ``baseline/_ref`` (the pip --target install that travels to the GPU
box, scripts/install_reference.sh) under a small in-memory compatibility layer
(``nums_b200.reference_compat``), because the reference pins ``numpy<=1.20`` and
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

from nums_b200 import reference_compat

REFERENCE_ROOT = reference_compat.reference_root() or "/root/reference"


def available() -> bool:
    return reference_compat.reference_root() is not None


def load():
    """Import the reference and return the ``nums`` package (serial system only).  The numpy-2 /
    no-ray shims live in ``nums_b200.reference_compat`` (they are also what lets the reference's host
    layers run over ``cuda_compute``)."""
    if not available():
        raise RuntimeError("reference not present (looked at $NUMS_REFERENCE_ROOT, /root/reference, baseline/_ref)")
    return reference_compat.load_reference("serial")


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
