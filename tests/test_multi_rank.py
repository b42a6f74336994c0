"""world_size-2 tests of the multi-GPU path: independent streams need no collective; the only exchange
step is the all-reduce of the 40 normal-equation sums in point-sharded tracking (BASELINE config 5)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(mode, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    import re
    res = [json.loads(m) for m in re.findall(r"RESULT (\{.*?\})(?=\s|RESULT|$)", r.stdout)]
    assert len(res) == 2
    return sorted(res, key=lambda d: d["rank"])


def test_shard_partition_is_disjoint_and_complete():
    from lsd_slam_b200 import sharded
    n = 160 * 120
    for world in (1, 2, 4, 8):
        parts = [sharded.shard_chunks(n, r, world) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 32


def test_gloo_world2_host_logic():
    res = _launch("cpu", 29631)
    for r in res:
        assert r["world"] == 2 and r["partition_ok"] and r["allreduce_ok"] and r["max_ok"]


@pytest.mark.gpu
def test_point_sharded_tracking_matches_single_gpu(seq_small, frames_small):
    """2 ranks (both on cuda:0, gloo all-reduce of the 40 sums) reproduce the unsharded host-LM tracker"""
    from lsd_slam_b200 import abi
    res = _launch("gpu", 29633)
    ctx = abi.Context(seq_small.w, seq_small.h, seq_small.K, max_frames=4)
    ctx.upload(0, frames_small[0][0])
    ctx.set_depth_gt(0, frames_small[0][1])
    ctx.upload(3, frames_small[3][0])
    trk = abi.SE3Tracker(ctx, mode=0)
    ref = trk.trackFrame(0, 3, np.array([0, 0, 0, 1, 0, 0, 0], np.float64))
    for r in res:
        p = np.array(r["pose"])
        assert np.linalg.norm(p[4:] - ref[4:]) <= 1e-4 * np.linalg.norm(ref[4:])
        assert np.abs(p[:4] - ref[:4]).max() <= 1e-6
    assert res[0]["pose"] == res[1]["pose"]          # ranks stay in lock step (identical decisions)
    ctx.close()
