"""multi-GPU paths, collected by pytest and skipped on boxes with fewer than 2 GPUs: launches the torchrun scripts of this directory
(one process per GPU, NCCL over NVLink) and checks their verdict"""
import os
import subprocess
import sys

import pytest

import ssq_testlib as T

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        return len([l for l in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.splitlines() if l.startswith("GPU ")])
    except OSError:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("script", ["dist_pipe_nccl.py", "dist_dupmark_nccl.py"])
def test_multi_gpu_scripts(tmp_path, script):
    n = min(_n_gpus(), 4) if script == "dist_pipe_nccl.py" else 2
    env = dict(os.environ, SSQ_DIST_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(T.ROOT, "tests", script)], capture_output=True, text=True, timeout=900, env=env, cwd=T.ROOT)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "identical_to_oracle" in r.stdout
