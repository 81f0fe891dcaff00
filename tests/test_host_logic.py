"""Host-side logic that needs no GPU: launch-group planning of the deferred flush, index handling of
the batched scatter."""
import numpy as np
import pytest


def test_plan_launch_groups_orders_by_readiness():
    from nums_b200.deferred import plan_launch_groups
    ev = object()
    # nothing in flight -> one launch, nothing to wait for
    assert plan_launch_groups([None, None, None], 0, 16) == [([0, 1, 2], None)]
    assert plan_launch_groups([(3, ev), (5, ev)], 5, 16) == [([0, 1], None)]
    # 8 x 8 blocked matmul streamed as: A row 0, B column by column, remaining A rows (bench.py e2e order)
    seq = {}
    n = 0
    for k in range(8):
        n += 1
        seq[("A", 0, k)] = n
    for j in range(8):
        for k in range(8):
            n += 1
            seq[("B", k, j)] = n
    for i in range(1, 8):
        for k in range(8):
            n += 1
            seq[("A", i, k)] = n
    needs = []
    for i in range(8):
        for j in range(8):
            latest = max(max(seq[("A", i, k)], seq[("B", k, j)]) for k in range(8))
            needs.append((latest, ev))
    plan = plan_launch_groups(needs, 0, 16)
    assert len(plan) == 15                                   # C(0, j) one by one, then one launch per block row
    assert [idx for idx, _ in plan[:8]] == [[j] for j in range(8)]
    assert [idx for idx, _ in plan[8:]] == [list(range(8 * i, 8 * i + 8)) for i in range(1, 8)]
    waits = [upto[0] for _, upto in plan]
    assert waits == sorted(waits) and waits[-1] == n        # every launch waits for more than the one before
    assert sorted(i for idx, _ in plan for i in idx) == list(range(64))
    # more distinct readiness levels than allowed launches -> neighbouring runs are merged
    many = [(s + 1, ev) for s in range(100)]
    plan = plan_launch_groups(many, 0, 16)
    assert len(plan) <= 16 and sorted(i for idx, _ in plan for i in idx) == list(range(100))
    assert all(upto[0] == max(idx) + 1 for idx, upto in plan)
    # already-awaited uploads do not force a wait
    plan = plan_launch_groups([(2, ev), None, (9, ev)], 4, 16)
    assert plan == [([0, 1], None), ([2], (9, ev))]


def test_ravel_indices_numpy_semantics():
    from nums_b200.cuda_compute import _ravel_indices
    rng = np.random.default_rng(3)
    shape = (4, 5, 6)
    ref = np.arange(int(np.prod(shape))).reshape(shape)
    idx = [tuple(int(rng.integers(-s, s)) for s in shape) for _ in range(200)]
    assert [int(v) for v in _ravel_indices(idx, shape)] == [int(ref[i]) for i in idx]
    assert [int(v) for v in _ravel_indices([(2,), (-1,)], (7,))] == [2, 6]
    for bad in [(4, 0, 0), (0, -6, 0), (0, 0, 6)]:
        with pytest.raises(IndexError):
            _ravel_indices([bad], shape)


def test_reference_system_subclasses_route_kernel_names_without_the_slow_getattribute():
    """reference_compat's system classes take object.__getattribute__ (the reference's Python-level override costs
    ~1 us per attribute access): kernel names must still resolve (through __getattr__) and must not collide with
    any real attribute of the classes."""
    from nums_b200 import reference_compat
    if not reference_compat.available():
        pytest.skip("reference not present")
    from oracle import np_oracle
    from oracle.cpu_system import OracleSystem
    cls = reference_compat.spmd_system_class()
    assert cls.__getattribute__ is object.__getattribute__
    assert reference_compat.cuda_system_class().__getattribute__ is object.__getattribute__
    system = cls(OracleSystem(), check=True)
    system.rng_cls = np_oracle.RNG
    system.init()
    from nums.core.systems.interfaces import ComputeInterface
    import inspect
    kernel_names = [n for n, _ in inspect.getmembers(ComputeInterface, predicate=inspect.isfunction)]
    assert len(kernel_names) >= 28
    # the reference's System inherits abstract stubs of the kernel names from ComputeInterface: the published
    # instance attributes must shadow them, and nothing else of the class may carry a kernel's name
    for klass in (cls, reference_compat.cuda_system_class()):
        for name in kernel_names + ["lr_grad_hess", "newton_step", "read_csv_block", "write_block_fs"]:
            owner = next((k for k in klass.__mro__ if name in k.__dict__), None)
            assert owner is None or owner.__name__ == "ComputeInterface", (name, owner)
    for name in kernel_names:
        assert getattr(system, name) is system.methods[name], name
    a = system.put(np.arange(6.0))
    out = system.bop("add", a, a, (6,), (6,), False, False, axes=None, syskwargs={"grid_entry": (0,), "grid_shape": (1,)})
    assert np.array_equal(system.get(out), 2 * np.arange(6.0))
