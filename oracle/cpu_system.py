"""TEST INFRASTRUCTURE ONLY -- a SerialSystem-shaped object over the CPU oracle.

Mirrors the reference's ``SerialSystem`` (systems.py:68-117: identity put/get, in-process calls
with ``syskwargs`` stripped) on top of ``oracle.np_oracle.OracleCompute``.  Used by the tests to
run the host drivers on CPU and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs to
time the reference's CPU path (NumPy / OpenBLAS on the host cores).  Never imported by the product.
"""
import numpy as np

from oracle.np_oracle import OracleCompute


class OracleSystem(object):

    def __init__(self, record=False):
        self.imp = OracleCompute()
        self.trace = [] if record else None
        self.registered = {}

    def get_block_addresses(self, grid):
        """SerialSystem.get_block_addresses (systems.py:131-141): everything lives on node:0."""
        return {entry: "node:0" for entry in grid.get_entry_iterator()}

    def register(self, name, func, remote_params=None):
        """System.register (systems.py:57-66): extra remote functions such as read_csv_block."""
        self.registered.setdefault(name, func)

    def put(self, value):
        return np.asarray(value)

    def get(self, oids):
        if isinstance(oids, list):
            return [self.get(o) for o in oids]
        return oids

    def call(self, name, *args, **kwargs):
        if self.trace is not None:
            from oracle.make_golden import call_signature, freeze
            self.trace.append(call_signature(name, freeze(args), freeze(kwargs)))
        kwargs = {k: v for k, v in kwargs.items() if k != "syskwargs"}
        if name in self.registered:
            return self.registered[name](*args, **kwargs)
        return getattr(self.imp, name)(*args, **kwargs)

    def __getattr__(self, name):
        if name.startswith("_") or not hasattr(OracleCompute, name):
            raise AttributeError(name)
        return lambda *a, **k: self.call(name, *a, **k)
