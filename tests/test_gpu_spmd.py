"""GPU-side parity of the multi-GPU plugin path (needs >= 2 GPUs; skipped on a 1-GPU box).

Launches scripts/spmd_gpu_check.py under torchrun at world size 2 (and 4 when available): the reference's
unmodified BlockArray / ArrayApplication / glms.newton over SpmdSystem(CudaSystem) on every rank, results
compared with NumPy on the device-fetched arrays at the BASELINE.json tolerances.
"""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_plugin_path_parity_multi_gpu(world):
    from nums_b200 import reference_compat
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    if not reference_compat.available():
        pytest.skip("reference not installed (scripts/install_reference.sh)")
    out = os.path.join(ROOT, "gpurun_out", "spmd_check_n%d.json" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "scripts", "spmd_gpu_check.py"), "--out", out]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-4000:]
    with open(out) as fh:
        record = json.load(fh)
    assert record["ok"] and record["world"] == world
    assert record["kernels_launched_rank0"] > 0
