"""GPU parity: per-frame staging kernels vs the oracle's Frame builders (bit-exact)."""
import numpy as np
import pytest

from lsd_slam_b200 import abi

pytestmark = pytest.mark.gpu


def test_pyramids_gradients_maxgrad_bit_exact(gpu_ctx_small, oracle, seq_small, frames_small):
    img, _ = frames_small[2]
    gpu_ctx_small.upload(7, img)
    of = oracle.Frame(7, img, seq_small.K)
    for lvl in range(5):
        assert np.array_equal(gpu_ctx_small.download(7, abi.BUF_IMAGE, lvl), of.image(lvl)), lvl
        assert np.array_equal(gpu_ctx_small.download(7, abi.BUF_GRADIENTS, lvl), of.gradients(lvl)), lvl
    assert np.array_equal(gpu_ctx_small.download(7, abi.BUF_MAXGRAD, 0), of.maxGradients(0))


def test_extreme_images(gpu_ctx_small, oracle, seq_small):
    rng = np.random.default_rng(5)
    for k, img in enumerate([np.zeros((240, 320), np.uint8), np.full((240, 320), 255, np.uint8),
                             rng.integers(0, 256, (240, 320), dtype=np.uint8)]):
        gpu_ctx_small.upload(20 + k, img)
        of = oracle.Frame(20 + k, img, seq_small.K)
        for lvl in range(5):
            assert np.array_equal(gpu_ctx_small.download(20 + k, abi.BUF_GRADIENTS, lvl), of.gradients(lvl))
        assert np.array_equal(gpu_ctx_small.download(20 + k, abi.BUF_MAXGRAD, 0), of.maxGradients(0))
        gpu_ctx_small.release(20 + k)


def test_gt_depth_import_and_idepth_pyramid_bit_exact(gpu_ctx_small, oracle, seq_small, frames_small):
    img, d = frames_small[0]
    d = d.copy()
    d[10:20, 10:20] = np.nan          # invalid GT depth cells (Frame.cpp:270)
    d[30:35, 40:50] = -1.0
    gpu_ctx_small.upload(0, img)
    gpu_ctx_small.set_depth_gt(0, d)
    of = oracle.Frame(0, img, seq_small.K)
    of.setDepthFromGroundTruth(d)
    for lvl in range(5):
        assert np.array_equal(gpu_ctx_small.download(0, abi.BUF_IDEPTH, lvl), of.idepth(lvl)), lvl
        assert np.array_equal(gpu_ctx_small.download(0, abi.BUF_IDEPTH_VAR, lvl), of.idepthVar(lvl)), lvl


def test_slot_exhaustion_and_errors(seq_small, frames_small):
    ctx = abi.Context(seq_small.w, seq_small.h, seq_small.K, max_frames=2)
    img, _ = frames_small[0]
    ctx.upload(1, img)
    ctx.upload(2, img)
    with pytest.raises(abi.LsdGpuError):
        ctx.upload(3, img)
    ctx.release(1)
    ctx.upload(3, img)
    with pytest.raises(abi.LsdGpuError):
        ctx.download(99, abi.BUF_IMAGE, 0)
    with pytest.raises(abi.LsdGpuError):
        abi.Context(100, 100, seq_small.K)            # not multiples of 16 (SlamSystem.cpp:55)
    ctx.close()
