"""The C++ host classes (lsd_slam_b200/host/lsd_host.h: SE3Tracker::trackFrame, DepthMap::updateKeyframe /
createKeyFrame with the reference's signatures) compile with g++ and, on the GPU, reproduce the Python/C-ABI loop."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_adapter_compiles_and_links():
    from lsd_slam_b200 import build
    build.build()
    build.build_host(force=True)
    assert os.path.exists(build.HOST_OUT) and os.path.exists(build.DEMO_OUT)
    syms = subprocess.run(["nm", "-DC", build.HOST_OUT], capture_output=True, text=True).stdout
    for s in ("lsd_slam::SE3Tracker::trackFrame(lsd_slam::TrackingReference*, lsd_slam::Frame*, lsd_slam::SE3 const&)",
              "lsd_slam::DepthMap::updateKeyframe(std::deque<std::shared_ptr<lsd_slam::Frame>",
              "lsd_slam::DepthMap::createKeyFrame(lsd_slam::Frame*)",
              "lsd_slam::Sim3Tracker::trackFrameSim3(lsd_slam::TrackingReference*, lsd_slam::Frame*, lsd_slam::Sim3 const&, int, int)",
              "lsd_slam::UndistorterPTAM::UndistorterPTAM(char const*)"):
        assert s in syms, s


@pytest.mark.gpu
def test_host_demo_matches_python_loop(tmp_path, seq_small, frames_small):
    from lsd_slam_b200 import abi, build
    from lsd_slam_b200.stream import GpuStream
    build.build_host()
    n = 9
    path = tmp_path / "frames.bin"
    with open(path, "wb") as f:
        np.array([seq_small.w, seq_small.h, n], np.int32).tofile(f)
        seq_small.K.astype(np.float32).tofile(f)
        frames_small[0][1].astype(np.float32).tofile(f)
        for k in range(n):
            frames_small[k][0].tofile(f)
    r = subprocess.run([build.DEMO_OUT, str(path), "5"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows = np.array([[float(x) for x in ln.split()] for ln in r.stdout.strip().splitlines()])
    ctx = abi.Context(seq_small.w, seq_small.h, seq_small.K, max_frames=8)
    gs = GpuStream(ctx, mode=1, kf_every=5)
    gs.init_gt(0, frames_small[0][0], frames_small[0][1])
    for k in range(1, n):
        gs.step(k, frames_small[k][0])
    py = np.array(gs.poses)
    assert rows.shape[0] == n - 1
    assert np.allclose(rows[:, 1:8], py, atol=1e-9), np.abs(rows[:, 1:8] - py).max()
    assert rows[:, 8].all()
    ctx.close()


def test_undistorter_config_file_parsing(tmp_path, oracle):
    """lsd_slam::UndistorterPTAM(configFileName) of the C++ adapter (util/Undistorter.cpp:100-166) -- host-only, no GPU --
    against the oracle's tables for the same four lines"""
    from lsd_slam_b200 import build
    build.build()
    build.build_host()
    fov = [0.535719308086809, 0.669566858850269, 0.493248545285398, 0.500408664348414, 0.897966326944875]
    for mode, out in (("crop", (640, 480)), ("full", (320, 240)), ("0.6 0.8 0.5 0.5 0", (320, 240))):
        cfg = tmp_path / "cam.cfg"
        cfg.write_text(" ".join(repr(v) for v in fov) + "\n640 480\n" + mode + "\n%d %d\n" % out)
        tables = tmp_path / "tables.bin"
        r = subprocess.run([build.DEMO_OUT, "--undistorter", str(cfg), str(tables)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        l1, l2 = r.stdout.strip().splitlines()
        assert [int(v) for v in l1.split()] == [1, 0, 640, 480, out[0], out[1]]
        oc = mode if mode in ("crop", "full") else [float(v) for v in mode.split()]
        want = oracle.UndistorterPTAM(fov, (640, 480), oc, out)
        K = np.array([float(v) for v in l2.split()], np.float32).reshape(3, 3)
        assert np.array_equal(K, want.K)
        t = np.fromfile(tables, np.float32).reshape(2, out[1], out[0])
        assert t[0].tobytes() == want.remapX.tobytes() and t[1].tobytes() == want.remapY.tobytes()
    # an OpenCV-model file (8 numbers on the first line) is not taken by the PTAM class
    cfg.write_text("500 500 320 240 0.1 0.01 0 0\n640 480\ncrop\n640 480\n")
    r = subprocess.run([build.DEMO_OUT, "--undistorter", str(cfg), str(tables)], capture_output=True, text=True, timeout=60)
    assert r.stdout.split()[0] == "0"
