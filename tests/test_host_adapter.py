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


def _qmul(a, b):
    x1, y1, z1, w1 = a
    x2, y2, z2, w2 = b
    return np.array([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                     w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])


def _qrot(q, v):
    qv = np.array([v[0], v[1], v[2], 0.0])
    qc = np.array([-q[0], -q[1], -q[2], q[3]])
    return _qmul(_qmul(q, qv), qc)[:3]


def _sim3_inv(a):
    qi = np.array([-a[0], -a[1], -a[2], a[3]])
    s = 1.0 / a[7]
    return np.concatenate([qi, s * _qrot(qi, -a[4:7]), [s]])


def _sim3_mul(a, b):
    q = _qmul(a[:4], b[:4])
    return np.concatenate([q / np.linalg.norm(q), a[4:7] + a[7] * _qrot(a[:4], b[4:7]), [a[7] * b[7]]])


@pytest.mark.gpu
def test_host_classes_map_a_frame_tracked_on_the_previous_keyframe(tmp_path, seq_small, frames_small):
    """DepthMap::updateKeyframe of the C++ adapter with a frame whose tracking parent is the OLD keyframe: the adapter chains
    FramePoseStruct::getCamToWorld (FramePoseStruct.cpp:84-105) into refToKf (DepthMap.cpp:1099) and hands it to
    lsdgpu_depth_update_keyframe_refs; the same loop through the Python mirror gives the same poses afterwards"""
    from lsd_slam_b200 import abi, build
    build.build_host()
    n, kf_every = 12, 5
    d0 = frames_small[0][1]
    depth0 = (d0 * float(np.mean(1.0 / d0[d0 > 0]))).astype(np.float32)   # mean inverse depth 1: no scale jump at the keyframe change
    path = tmp_path / "frames.bin"
    with open(path, "wb") as f:
        np.array([seq_small.w, seq_small.h, n], np.int32).tofile(f)
        seq_small.K.astype(np.float32).tofile(f)
        depth0.tofile(f)
        for k in range(n):
            frames_small[k][0].tofile(f)
    rows = {}
    for late in (False, True):
        r = subprocess.run([build.DEMO_OUT, str(path), str(kf_every)] + (["late"] if late else []), capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        rows[late] = np.array([[float(x) for x in ln.split()] for ln in r.stdout.strip().splitlines()])
    assert np.abs(rows[True][6:, 1:8] - rows[False][6:, 1:8]).max() > 1e-7          # the extra mapping step changed the map
    ctx = abi.Context(seq_small.w, seq_small.h, seq_small.K, max_frames=12)
    trk, dm = abi.SE3Tracker(ctx, mode=1), abi.DepthMap(ctx)
    trk.settings.maxItsPerLvl[4] = 0
    ctx.upload(0, frames_small[0][0])
    ctx.set_depth_gt(0, depth0)
    dm.initializeFromGTDepth(0)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    kf, early, last, poses = 0, None, ident, []
    cam_to_world = {0: np.array([0, 0, 0, 1, 0, 0, 0, 1.0])}
    for k in range(1, n):
        ctx.upload(k, frames_small[k][0])
        if ctx.depth_updated_flag(kf):
            trk.importFrame(kf)
        pose = np.array(trk.trackFrame(kf, k, last))
        poses.append(pose)
        cam_to_world[k] = _sim3_mul(cam_to_world[kf], np.concatenate([pose, [1.0]]))
        if k % kf_every == 0:
            dm.finalizeKeyFrame()
            q = dm.createKeyFrame(k)
            cam_to_world[k] = _sim3_mul(cam_to_world[kf], q)
            kf, last = k, ident
            if early is not None:
                dm.updateKeyframe([(early, _sim3_mul(_sim3_inv(cam_to_world[kf]), cam_to_world[early]))])
            early = None
        else:
            dm.updateKeyframe([k])
            ctx.clear_good_mask(k)
            last = pose
            if early is None:
                early = k
    ctx.close()
    assert np.allclose(rows[True][:, 1:8], np.array(poses), atol=1e-9), np.abs(rows[True][:, 1:8] - np.array(poses)).max()


def test_host_classes_chain_the_absolute_poses_like_the_reference(tmp_path, oracle):
    """FramePoseStruct::getCamToWorld (FramePoseStruct.cpp:84-105) and refToKf = kf.camToWorld^-1 * frame.camToWorld
    (DepthMap.cpp:1099) of the C++ adapter against the oracle's Sim3 product / inverse -- host code only, no GPU"""
    import ctypes as C
    from lsd_slam_b200 import build
    build.build()
    build.build_host()
    exe = tmp_path / "host_sim3_chain"
    here = os.path.dirname(build.HOST_OUT)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "native", "host_sim3_chain.cpp"),
                        "-L" + here, "-Wl,-rpath," + here, "-l:liblsdgpu_host.so", "-l:liblsdgpu.so"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    got, kf1, fr = (np.array([float(x) for x in ln.split()]) for ln in out.stdout.strip().splitlines())
    dp = C.POINTER(C.c_double)

    def call(fn, *args):
        o = np.zeros(8)
        arrs = [np.ascontiguousarray(a, np.float64) for a in args]
        getattr(oracle.lib(), fn)(*[a.ctypes.data_as(dp) for a in arrs], o.ctypes.data_as(dp))
        return o
    want = call("lsdo_sim3d_mul", call("lsdo_sim3d_inverse", kf1), fr)
    assert np.abs(got - want).max() <= 1e-14, got - want


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
