"""The host drivers in nums_b200.blocks issue the SAME per-block kernel calls as the reference.

For each scenario of oracle/scenarios.py the reference's ordered call signatures (kernel name,
argument shapes/dtypes/flags, syskwargs placement hints) and final results were recorded from the
real reference by oracle/make_golden.py.  Here the mirror runs over an oracle-backed system on
CPU and must reproduce the sequence call for call and the results bit for bit.
"""
import numpy as np
import pytest

from oracle import scenarios
from tests.helpers import OracleSystem, canon_r, load_golden, rel_fro


@pytest.fixture(scope="module")
def golden():
    return load_golden("ref_scenarios.pkl.gz")


@pytest.mark.parametrize("name", sorted(scenarios.SCENARIOS))
def test_same_kernel_call_sequence_and_results(golden, name):
    system = OracleSystem()
    api = scenarios.MirrorApi(system)
    system.trace = []
    results = scenarios.SCENARIOS[name](api)
    want = golden[name]
    drop = ("touch",)
    mine = [s for s in system.trace if s[0] not in drop]
    ref = [s for s in want["signatures"] if s[0] not in drop]
    if name not in scenarios.RESULTS_ONLY:
        for i, (a, b) in enumerate(zip(mine, ref)):
            assert a == b, "call %d differs:\n mirror   : %r\n reference: %r" % (i, a, b)
        assert len(mine) == len(ref), (len(mine), len(ref))
    for key, value in want["results"].items():
        got = np.asarray(results[key])
        assert got.shape == value.shape and got.dtype == value.dtype, key
        assert np.array_equal(got, value, equal_nan=True), key
