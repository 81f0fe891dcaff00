"""conftest.py for the reference's own test-suite (copied over baseline/_ref/reference_tests/conftest.py
by scripts/install_reference.sh / scripts/run_reference_suite.py).

Same fixtures as /root/reference/tests/conftest.py:28-72 -- ``app_inst`` (module scoped, built by
``get_app(mode)``) and ``nps_app_inst`` (the global application) -- with the ray modes, which need a
ray cluster, replaced by the one new mode a maintainer would add: ``"cuda"`` = the reference's
ArrayApplication over ``CudaSystem`` + ``cuda_compute`` (nums_b200.reference_compat).  Modes are
chosen with ``NUMS_TEST_MODES`` (comma separated, default ``serial,cuda``).
"""
import os
import sys

import pytest

_REPO = os.environ.get("NUMS_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from nums_b200 import reference_compat  # noqa: E402

reference_compat.load_reference()

from nums.core.systems import numpy_compute  # noqa: E402
from nums.core.systems.systems import SerialSystem  # noqa: E402
from nums.core.systems.filesystem import FileSystem  # noqa: E402
from nums.core.array.application import ArrayApplication  # noqa: E402

MODES = [m for m in os.environ.get("NUMS_TEST_MODES", "serial,cuda").split(",") if m]


@pytest.fixture(scope="module", params=MODES)
def app_inst(request):
    app = get_app(request.param)
    yield app
    app.system.shutdown()


@pytest.fixture(scope="module", params=MODES)
def nps_app_inst(request):
    from nums.core import settings
    from nums.core import application_manager
    settings.system_name = request.param     # "cuda" is registered by nums_b200.reference_compat
    yield application_manager.instance()
    application_manager.destroy()


def get_app(mode):
    if mode == "serial":
        system = SerialSystem(compute_module=numpy_compute)
    elif mode == "cuda":
        return reference_compat.cuda_app()
    else:
        raise Exception("unknown mode %s" % mode)
    system.init()
    return ArrayApplication(system=system, filesystem=FileSystem(system))
