"""Multi-rank runs of the CUDA path: ``sample()`` sharded over W ranks must return particles BIT-IDENTICAL to the
single-rank run (every reduction order in the step is a function of (M, D) only -- DESIGN.md section 5).

Two set-ups, both spawned with ``torch.distributed.run`` like bench.py's N > 1 launch:
  * W ranks on W GPUs over NCCL + peer memory (needs >= 2 devices: runs in the driver's scaling lease);
  * 2 ranks time-sliced on ONE GPU with a gloo process group: the peer-memory exchange (CUDA IPC, remote stores,
    flag waits: kernels_peer.cuh) runs end to end on a single-GPU box too, and so does the NCCL-free fallback logic.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "mgpu_check.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world, out, extra_env, timeout=420):
    env = dict(os.environ, MGPU_OUT=out, MGPU_WATCHDOG="180", DIBS_B200_PEER_TIMEOUT_MS="20000", **extra_env)
    if world == 1:
        cmd = [sys.executable, TOOL]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), TOOL]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]


def _compare(out, world, cases):
    for name in cases:
        a = np.load(os.path.join(out, f"mgpu_w1_{name}.npz"))
        b = np.load(os.path.join(out, f"mgpu_w{world}_{name}.npz"))
        assert np.array_equal(a["z"], b["z"]), (name, float(np.abs(a["z"] - b["z"]).max()))
        assert np.array_equal(a["theta"], b["theta"]), name


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 8])
def test_sample_bit_identical_across_gpus(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path)
    _run(1, out, {})
    _run(world, out, {"MGPU_EXPECT": "peer-memory"})
    _compare(out, world, ("bge", "lin", "nn", "linL"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sample_bit_identical_nccl_fallback(tmp_path):
    out = str(tmp_path)
    _run(1, out, {"MGPU_CASES": "lin"})
    _run(2, out, {"MGPU_CASES": "lin", "DIBS_B200_NO_P2P": "1", "MGPU_EXPECT": "nccl"})
    _compare(out, 2, ("lin",))


def test_two_ranks_one_gpu_peer_memory(tmp_path):
    """2 ranks sharing cuda:0: the sharded step over CUDA-IPC peer memory vs the single-rank run, bit for bit."""
    out = str(tmp_path)
    _run(1, out, {"MGPU_CASES": "lin,bge,linL"})
    _run(2, out, {"MGPU_CASES": "lin,bge,linL", "MGPU_SAME_GPU": "1", "MGPU_EXPECT": "peer-memory"})
    _compare(out, 2, ("lin", "bge", "linL"))
