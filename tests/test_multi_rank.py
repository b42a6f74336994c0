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


def _tp_chunk_of(b, k, G):
    """csrc/track_persistent.cuh tpChunkOf"""
    return k * G + (b + k * 37) % G


def test_virtual_grid_of_the_peer_sharded_tracker_owns_every_chunk_exactly_once():
    """config 5, in-kernel exchange: the CTAs of all ranks form one virtual grid (CTA b of rank r is number b * N + r of G * N);
    slot k of virtual CTA v owns chunk tpChunkOf(v, k, G * N) if it exists.  Every 32-pixel chunk of every tracked level must have
    exactly one owner, for every number of ranks, and the ranks' shares must be balanced."""
    G = 148
    for (w, h) in ((640, 480), (1280, 1024), (320, 240)):
        for lvl in (1, 2, 3, 4):
            n_int = ((w >> lvl) - 2) * ((h >> lvl) - 2)
            n_chunks = (n_int + 31) // 32
            for N in (1, 2, 4, 8):
                Gv = G * N
                n_slots = (n_chunks + Gv - 1) // Gv
                owner = {}
                per_rank = [0] * N
                for r in range(N):
                    for b in range(G):
                        v = b * N + r
                        for k in range(n_slots):
                            c = _tp_chunk_of(v, k, Gv)
                            if c < n_chunks:
                                assert c not in owner, (w, h, lvl, N, c)
                                owner[c] = (r, b, k)
                                per_rank[r] += 1
                assert len(owner) == n_chunks
                assert max(per_rank) - min(per_rank) <= max(1, n_chunks // (8 * N)) + G, (w, h, lvl, N, per_rank)


def _attach_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lsd_slam_b200 import sharded

    class FakeCtx:
        def __init__(self):
            self.attached = None

        def peer_export(self):
            return bytes([rank]) * 64

        def peer_attach(self, r, handles):
            self.attached = (r, handles)

        def peer_detach(self):
            self.attached = None
    ctx = FakeCtx()
    sharded.attach_peers(ctx, dist)
    r, handles = ctx.attached
    ok = (r == rank and len(handles) == world and all(handles[i] == bytes([i]) * 64 for i in range(world)))
    sharded.detach_peers(ctx, dist)
    q.put((rank, ok and ctx.attached is None))
    dist.destroy_process_group()


def test_peer_handles_are_gathered_in_rank_order_gloo_world_2():
    """host side of lsdgpu_peer_attach: every rank receives all ranks' IPC handles in rank order (torch.distributed, gloo here)"""
    import multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    ps = [ctxm.Process(target=_attach_worker, args=(r, 2, 29541, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert out == [(0, True), (1, True)]
