"""Runs the reference's own tests (baseline/_ref/reference_tests, see scripts/install_reference.sh) in one
mode and writes ``{nodeid: outcome}`` as JSON.

    python scripts/run_reference_suite.py --mode cuda   --out gpurun_out/ref_suite_cuda.json
    python scripts/run_reference_suite.py --mode serial --out tests/golden/ref_suite_serial.json

``serial`` = the reference's SerialSystem + numpy_compute (the pass-set that NumPy 2.3 allows);
``cuda`` = the same host layers over CudaSystem + cuda_compute.  tests/test_gpu_reference_suite.py asserts
that every test passing in ``serial`` passes in ``cuda``.
"""
import argparse
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(REPO, "baseline", "_ref", "reference_tests")

# the hot-path suites; storage/S3/app-manager/api tests need moto / ray
DEFAULT_TARGETS = ["core/array", "numpy", "models"]


class Recorder(object):
    def __init__(self):
        self.outcomes = {}
        self.details = {}

    def pytest_runtest_logreport(self, report):
        nodeid = report.nodeid
        if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
            outcome = report.outcome
            if report.when == "setup" and outcome == "failed":
                outcome = "error"
            self.outcomes[nodeid] = outcome
            if outcome in ("failed", "error"):
                self.details[nodeid] = str(report.longrepr)[-1500:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cuda")
    ap.add_argument("--out", required=True)
    ap.add_argument("--details", default=None)
    ap.add_argument("targets", nargs="*")
    args, extra = ap.parse_known_args()
    if not os.path.isdir(SUITE):
        print("reference tests not installed (run scripts/install_reference.sh)", file=sys.stderr)
        return 3
    shutil.copyfile(os.path.join(REPO, "scripts", "ref_conftest.py"), os.path.join(SUITE, "conftest.py"))
    os.environ["NUMS_TEST_MODES"] = args.mode
    os.environ["NUMS_B200_ROOT"] = REPO
    os.environ.setdefault("NUMS_REFERENCE_ROOT", os.path.join(REPO, "baseline", "_ref"))
    import pytest
    rec = Recorder()
    targets = [os.path.join(SUITE, t) for t in (args.targets or DEFAULT_TARGETS)]
    args.out = os.path.abspath(args.out)
    args.details = os.path.abspath(args.details) if args.details else None
    os.chdir(SUITE)
    rc = pytest.main(["--rootdir", SUITE, "--confcutdir", SUITE, "-q", "-p", "no:cacheprovider", "--tb=short"] + extra + targets, plugins=[rec])
    strip = SUITE + os.sep
    out = {k.replace(strip, ""): v for k, v in sorted(rec.outcomes.items())}
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    if args.details:
        with open(args.details, "w") as fh:
            json.dump({k.replace(strip, ""): v for k, v in rec.details.items()}, fh, indent=1, sort_keys=True)
    counts = {}
    for v in out.values():
        counts[v] = counts.get(v, 0) + 1
    print("reference suite [%s]: %s (pytest rc %s)" % (args.mode, counts, rc))
    return 0


if __name__ == "__main__":
    sys.exit(main())
