"""UndistorterPTAM (util/Undistorter.cpp:91-411): the host-side table builder of the library against the
oracle (CPU, bit-exact), and -- on the GPU -- undistort() and the fused undistort + Frame construction against the oracle."""
import numpy as np
import pytest

from lsd_slam_b200 import abi

# a wide-angle PTAM/ATAN calibration (relative units) as in the reference's example files
FOV = [0.535719308086809, 0.669566858850269, 0.493248545285398, 0.500408664348414, 0.897966326944875]
CASES = [
    (FOV, (640, 480), "crop", (640, 480)),
    (FOV, (752, 480), "full", (640, 480)),
    (FOV, (640, 480), (0.6, 0.8, 0.5, 0.5, 0), (320, 240)),
    ([0.5, 0.7, 0.49, 0.51, 0.0], (640, 480), "crop", (640, 480)),      # no distortion, same size: pass-through
    ([0.5, 0.7, 0.49, 0.51, 0.0], (640, 480), "crop", (320, 240)),      # no distortion, resize only
]


@pytest.mark.parametrize("case", CASES)
def test_prepare_matches_oracle_bit_exact(oracle, case):
    lib_u = abi.UndistorterPTAM(*case)
    ora_u = oracle.UndistorterPTAM(*case)
    assert lib_u.status == ora_u.status
    assert lib_u.K.tobytes() == ora_u.K.tobytes()
    assert lib_u.remapX.tobytes() == ora_u.remapX.tobytes()
    assert lib_u.remapY.tobytes() == ora_u.remapY.tobytes()
    assert lib_u.K[2, 2] == 1 and lib_u.K[0, 0] > 0 and lib_u.K[1, 1] > 0


def test_tables_are_sane(oracle):
    u = oracle.UndistorterPTAM(*CASES[0])
    assert u.status == 0 and (u.remapX >= 0).all()                      # "crop": every output pixel has a source
    cx, cy = u.K[0, 2], u.K[1, 2]
    # the principal point maps onto the input principal point (no distortion at r = 0)
    x0, y0 = int(round(cx)), int(round(cy))
    assert abs(u.remapX[y0, x0] - (FOV[2] * 640 - 0.5)) < 1.5 and abs(u.remapY[y0, x0] - (FOV[3] * 480 - 0.5)) < 1.5
    full = oracle.UndistorterPTAM(*CASES[1])
    assert (full.remapX < 0).any() and (full.remapX[full.remapX < 0] == -1).all()      # "full": black corners
    ident = oracle.UndistorterPTAM(*CASES[3])
    assert ident.status == 1
    img = np.random.default_rng(0).integers(0, 256, (480, 640)).astype(np.uint8)
    assert np.array_equal(ident.undistort(img), img)
    # a constant image stays constant wherever a source exists
    flat = np.full((480, 752), 77, np.uint8)
    out = full.undistort(flat)
    assert set(np.unique(out)) <= {0, 76, 77}


def test_table_validation(oracle):
    """host-only guard of lsdgpu_set_undistorter: the prepared tables pass, corrupted ones are rejected"""
    import ctypes as C
    L = abi.load()
    fp = C.POINTER(C.c_float)
    for case in CASES:
        u = oracle.UndistorterPTAM(*case)
        (iw, ih), (ow, oh) = case[1], case[3]
        assert L.lsdgpu_undistorter_validate_tables(iw, ih, ow, oh, u.remapX.ctypes.data_as(fp), u.remapY.ctypes.data_as(fp)) == 0
    u = oracle.UndistorterPTAM(*CASES[0])
    (iw, ih), (ow, oh) = CASES[0][1], CASES[0][3]
    for bad_x, bad_y in ((iw - 1.0, 5.0), (5.0, ih - 1.0), (float("nan"), 5.0), (5.0, -0.5), (1e9, 1e9)):
        x, y = u.remapX.copy(), u.remapY.copy()
        x[7, 9], y[7, 9] = bad_x, bad_y
        assert L.lsdgpu_undistorter_validate_tables(iw, ih, ow, oh, x.ctypes.data_as(fp), y.ctypes.data_as(fp)) == 1 + 7 * ow + 9
    assert L.lsdgpu_undistorter_validate_tables(iw, ih, ow, oh, None, None) == -2


def _raw(seq_w, seq_h, seed=3):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (seq_h // 8 + 2, seq_w // 8 + 2)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8), np.float32))[:seq_h, :seq_w]
    img += rng.normal(0, 6, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_undistort_and_fused_frame_bit_exact(oracle, case):
    lib_u, ora_u = abi.UndistorterPTAM(*case), oracle.UndistorterPTAM(*case)
    ow, oh = case[3]
    ctx = abi.Context(ow, oh, lib_u.getK(), max_frames=4)
    lib_u.install(ctx)
    raw = _raw(*case[1])
    want = ora_u.undistort(raw)
    assert np.array_equal(ctx.undistort(raw), want)
    ctx.upload_distorted(7, raw)
    of = oracle.Frame(7, want, ora_u.K)
    for lvl in range(5):
        assert np.array_equal(ctx.download(7, abi.BUF_IMAGE, lvl), of.image(lvl)), lvl
    assert np.array_equal(ctx.download(7, abi.BUF_GRADIENTS, 0)[..., :3], of.gradients(0)[..., :3])
    assert np.array_equal(ctx.download(7, abi.BUF_MAXGRAD, 0), of.maxGradients(0))
    # the undistorted path and the plain path agree on the same undistorted image
    ctx.upload(8, want)
    assert np.array_equal(ctx.download(8, abi.BUF_IMAGE, 2), ctx.download(7, abi.BUF_IMAGE, 2))
    ctx.close()


@pytest.mark.gpu
def test_upload_distorted_needs_undistorter(seq_small):
    ctx = abi.Context(seq_small.w, seq_small.h, seq_small.K, max_frames=4)
    with pytest.raises(abi.LsdGpuError):
        ctx.upload_distorted(0, np.zeros((seq_small.h, seq_small.w), np.uint8))
    ctx.close()
