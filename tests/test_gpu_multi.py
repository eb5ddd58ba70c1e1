"""Multi-GPU checks on real hardware (skipped with fewer than two visible GPUs): one process per
GPU under torchrun, NCCL backend.  The host-side logic is covered on CPU by test_parallel_gloo."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def test_sharded_runs_equal_single_engine_runs_over_nccl(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(HERE, "multigpu_worker.py"), str(tmp_path)]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    msgs = []
    for r in range(2):
        p = tmp_path / ("rank%d.txt" % r)
        msgs.append(p.read_text() if p.exists() else "no result file")
    assert proc.returncode == 0 and all(m == "ok" for m in msgs), "\n".join(msgs) + proc.stderr[-2000:]
