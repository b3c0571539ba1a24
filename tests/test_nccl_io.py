"""Distributed encode + file gather over NCCL (BASELINE configs[4] at a reduced stream count): one process per GPU,
all-gather of the byte counts, serial-writer gather to rank 0 (io_common.write_compressed), read-back of a keep +
stream_slice window checked against the CPU oracle.  Needs two GPUs (NCCL refuses two ranks on one device); the
world-size-2 host logic is covered on CPU by tests/test_mpi_gloo.py and tests/test_io_layout.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg5_two_ranks_nccl():
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "full_configs.py"), "5", "--scale", "0.0024"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert lines, res.stdout[-2000:] + res.stderr[-2000:]
    r = json.loads(lines[-1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "cfg5_2gpu_scaled.json"), "w") as fh:
        fh.write(lines[-1] + "\n")
    assert r["cfg"] == 5 and r["n_gpus"] == 2
    assert r["ok"] and r["oracle_ok"] and r["ok_idx"], r
    assert 0.3 < r["ratio"] < 0.7
