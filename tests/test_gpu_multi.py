"""GPU: the shared-dof exchange + all-reduce dots against the serial solve, on both transports.

  * >= 2 devices (skipped on a 1-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
    one rank per GPU over the peer-memory path (default) and over NCCL send/recv + all-reduce (B200PA_NO_P2P=1);
  * ONE device: two (and four) ranks that share cuda:0 - the peer-memory kernels (k_px_send / k_px_recv / k_px_allreduce,
    csrc/comm.cu) run exactly as between GPUs, the mailboxes are CUDA-IPC mappings between the processes; NCCL refuses
    ranks on one device, so the communicator is created without it (nccl_id = NULL)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_worker(world, p, port, env_extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "mp_gpu_worker.py"), str(p)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env_extra))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "MULTI_OK" in out.stdout
    return out.stdout


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("world,p", [(2, 2), (2, 3), (4, 2), (8, 1)])
def test_partitioned_gpu_matches_serial(world, p, transport):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = run_worker(world, p, 29700 + world * 10 + p + (100 if transport == "nccl" else 0),
                     {"B200PA_NO_P2P": "1" if transport == "nccl" else "0"})
    assert f"p2p={transport == 'p2p'}" in out


@pytest.mark.parametrize("world,p", [(2, 2), (2, 4), (4, 3)])
def test_ranks_sharing_one_gpu_match_serial(world, p):
    out = run_worker(world, p, 29900 + world * 10 + p, {"B200PA_TEST_ONE_GPU": "1", "B200PA_NO_P2P": "0"})
    assert "p2p=True" in out
