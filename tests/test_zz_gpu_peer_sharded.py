"""BASELINE config 5 through the in-kernel peer exchange (include/lsdgpu.h "one stream over several GPUs"): two processes map each
other's arenas with CUDA IPC and the device-resident tracker shards every level over them, exchanging the per-pass sums and the
level-1 mask through the mapped memory.  A one-GPU box runs both ranks on cuda:0 (the driver time-slices the two cooperative
kernels -- slow, but the protocol, the chunk split, the rank-ordered sums and the mask hand-over are the ones the multi-GPU runs
use; `profiles/r2b_peer_sharded_{2,4}gpu.json` are those runs)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_shared_between_processes():
    """two processes can only share cuda:0 in the DEFAULT compute mode (not EXCLUSIVE_PROCESS / PROHIBITED)"""
    try:
        import pynvml
        pynvml.nvmlInit()
        mode = pynvml.nvmlDeviceGetComputeMode(pynvml.nvmlDeviceGetHandleByIndex(0))
        return mode == pynvml.NVML_COMPUTEMODE_DEFAULT
    except Exception:
        return True


@pytest.mark.gpu
def test_two_ranks_track_one_stream_and_stay_bit_identical():
    # (file name: runs last in the suite -- it is the only test that needs two processes on one device)
    if not _gpu_shared_between_processes():
        pytest.skip("GPU 0 is not in the default compute mode: two processes cannot share it")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "bench_peer_sharded.py"), "--frames", "7", "--kf-every", "4",
           "--width", "320", "--height", "240", "--same-gpu"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["all_ranks_bit_identical"] is True
    # against the same loop on one GPU: only the summation order of the 40 sums differs (two per-GPU totals added instead of one)
    assert d["max_pose_abs_diff"] <= 2e-6
    assert d["n_valid"] > 20000
