"""GPU test of the sharded multiply (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("algo", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_multiply_nccl(world, algo):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MATREL_GEMM_ALGO=str(algo))
    port = 29700 + world * 5 + algo + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    if r.returncode != 0:   # the workers' own tracebacks, not torchrun's summary of them
        lines = [ln for ln in r.stderr.splitlines() if "[rank" in ln or "Error" in ln or "assert" in ln]
        raise AssertionError(r.stdout[-1500:] + "\n".join(lines[:60]))
    assert f"OK world={world}" in r.stdout
