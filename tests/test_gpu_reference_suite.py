"""The reference's OWN test-suite as the acceptance suite of the backend (north star: "cuda_compute
passes the repo's tests/core and tests/numpy suites").

``scripts/install_reference.sh`` installs the unmodified reference and its tests into the git-ignored
``baseline/_ref`` (which travels to the GPU box).  ``scripts/run_reference_suite.py --mode cuda`` runs
``tests/core/array``, ``tests/numpy`` and ``tests/models`` of the reference with ``app_inst`` /
``nps_app_inst`` = the reference's ArrayApplication over ``CudaSystem`` + ``cuda_compute``
(scripts/ref_conftest.py: the reference's ``get_app`` with a ``"cuda"`` mode).  Every test that
passes with the reference's SerialSystem + numpy_compute on this NumPy
(``tests/golden/ref_suite_serial.json``, produced by the same runner with ``--mode serial``; the 11
serial failures are the fork / NumPy-2 defects listed in SURVEY.md section 4) must pass on the GPU,
one pytest item per reference test.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "baseline", "_ref", "reference_tests")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_suite_serial.json")

with open(GOLDEN) as _fh:
    SERIAL = json.load(_fh)
SERIAL_PASSED = sorted(k for k, v in SERIAL.items() if v == "passed")


def _cuda_id(serial_id):
    """Tests that take app_inst / nps_app_inst carry the mode in their id; the others (pure host index
    algebra in test_selection.py / test_broadcasting.py, test_explicit_init) have the same id in both runs."""
    if serial_id.endswith("serial]"):
        return serial_id[:-len("serial]")] + "cuda]"
    return serial_id


@pytest.fixture(scope="session")
def cuda_outcomes():
    if not os.path.isdir(SUITE):
        pytest.skip("reference not installed under baseline/_ref (scripts/install_reference.sh)")
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "ref_suite_cuda.json")
    details = os.path.join(out_dir, "ref_suite_cuda_details.json")
    for p in (out, details):
        if os.path.exists(p):
            os.remove(p)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_reference_suite.py"), "--mode", "cuda",
                           "--out", out, "--details", details],
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
    with open(os.path.join(out_dir, "ref_suite_cuda.log"), "w") as fh:
        fh.write(proc.stdout)
    assert os.path.exists(out), proc.stdout[-3000:]
    with open(out) as fh:
        outcomes = json.load(fh)
    det = {}
    if os.path.exists(details):
        with open(details) as fh:
            det = json.load(fh)
    return outcomes, det


@pytest.mark.gpu
def test_reference_suite_pass_set_matches_serial(cuda_outcomes):
    outcomes, _det = cuda_outcomes
    cuda_passed = {k for k, v in outcomes.items() if v == "passed"}
    missing = [k for k in SERIAL_PASSED if _cuda_id(k) not in cuda_passed]
    assert not missing, "pass with numpy_compute but not with cuda_compute: %s" % missing
    assert len(SERIAL_PASSED) >= 81


@pytest.mark.gpu
@pytest.mark.parametrize("serial_id", SERIAL_PASSED)
def test_reference_test_on_gpu(cuda_outcomes, serial_id):
    outcomes, det = cuda_outcomes
    cid = _cuda_id(serial_id)
    assert outcomes.get(cid) == "passed", "%s: %s\n%s" % (cid, outcomes.get(cid), det.get(cid, ""))
