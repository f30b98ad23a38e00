"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped on a one-GPU box): the marker-sharded pipeline under torchrun
with NCCL must reproduce the one-GPU results -- integer cross-products and kinship bit-identical, scan outputs within
1e-13.  The CPU-side plumbing of the same path is covered by tests/test_dist_cpu.py (gloo)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_sharded_pipeline_matches_one_gpu():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if ngpu < 4 else 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=580)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] and rep["gram_bit_identical"] and rep["kinship_bit_identical"] and rep["scan_max_rel_diff"] <= 1e-13
