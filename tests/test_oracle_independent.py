"""A SECOND, independent restatement of parts of the reference in vectorised numpy (fp32 element ops are IEEE, no FMA),
written from the reference text and not from oracle/*.c, to cross-check the C oracle where the reference itself cannot
be run (parity is unpinned):

  * regularizeDepthMapRow<occl>         DepthEstimation/DepthMap.cpp:758-848    -> bit-exact against the oracle
  * regularizeDepthMapFillHolesRow      DepthEstimation/DepthMap.cpp:656-701    -> bit-exact
  * buildRegIntegralBuffer              DepthEstimation/DepthMap.cpp:724-756    -> exact (integers)
  * SE3Tracker one evaluation           Tracking/SE3Tracker.cpp:885-1029, 749-790, 1258-1299 + LGS6 -> sums to fp32 summation-order accuracy
  * Sim3Tracker one evaluation          Tracking/Sim3Tracker.cpp:414-607, 748-856, 992-1047 + LGS4/6/7 -> same
  * TrackingReference::makePointCloud   Tracking/TrackingReference.cpp:96-147   -> through the two evaluations
  * observeDepthRow / Create / Update, makeAndCheckEPL, doLineStereo, prepareForStereoWith (tests/restate_stereo.py,
    pure Python with fp32 scalars, sampled pixels)                              -> bit-exact
  * propagateDepth (tests/restate_stereo.py, all pixels, both admission tests)   DepthMap.cpp:475-653  -> bit-exact
  * Frame::buildMaxGradients, setDepth, buildIDepthAndIDepthVar                  Frame.cpp:690-767, 199-243, 775-877 -> bit-exact
  * DepthMap::createKeyFrame (:1222-1327) composed from the pieces above, rescale included        -> bit-exact
  * the LM loops of SE3Tracker::trackFrame (:280-486) and Sim3Tracker::trackFrameSim3 (:149-382), driven from Python over
    the oracle's single evaluations: identical accept / reject / lambda schedules and call counts, identical poses
"""
import numpy as np
import pytest

from lsd_slam_b200 import synth

F = np.float32


def unzero(v):
    """UNZERO, util/settings.h:34: double literals, result stored to float"""
    d = v.astype(np.float64)
    out = np.where(d < 0, np.where(d > -1e-10, -1e-10, d), np.where(d < 1e-10, 1e-10, d))
    return out.astype(F)


def _pad(a, r, fill=0):
    return np.pad(a, r, constant_values=fill)


def np_integral(hyp):
    v = np.where(hyp["isValid"] > 0, hyp["validity_counter"], 0).astype(np.int32)
    return np.cumsum(np.cumsum(v, axis=1, dtype=np.int32), axis=0, dtype=np.int32)


def np_fill_holes(hyp, maxgrad, min_use_grad=5.0):
    H, W = hyp.shape
    out = hyp.copy()
    valid = hyp["isValid"] > 0
    val = np.where(valid, hyp["validity_counter"], 0).astype(np.int32)
    pv, pvalid = _pad(val, 2), _pad(valid, 2, False)
    pid, pvar = _pad(hyp["idepth"], 2), _pad(np.where(valid, hyp["idepth_var"], F(1)), 2, F(1))
    vsum = np.zeros((H, W), np.int32)
    sumId = np.zeros((H, W), F)
    sumIv = np.zeros((H, W), F)
    for dy in range(-2, 3):              # rows outer, columns inner: the order of the source loops (:679-688)
        for dx in range(-2, 3):
            sl = (slice(2 + dy, 2 + dy + H), slice(2 + dx, 2 + dx + W))
            vsum += pv[sl]
            m = pvalid[sl]
            sumId = np.where(m, sumId + pid[sl] / pvar[sl], sumId)
            sumIv = np.where(m, sumIv + F(1) / pvar[sl], sumIv)
    yy, xx = np.mgrid[0:H, 0:W]
    region = (xx >= 3) & (xx < W - 2) & (yy >= 3) & (yy < H - 2)
    cand = region & ~valid & ~(maxgrad < F(min_use_grad))
    create = cand & (((hyp["blacklisted"] >= -1) & (vsum > 30)) | (vsum > 100))        # MIN_BLACKLIST, VAL_SUM_MIN_FOR_CREATE / _UNBLACKLIST
    with np.errstate(all="ignore"):
        obs = unzero(sumId / sumIv)
    out["isValid"][create] = 1
    out["blacklisted"][create] = 0
    out["validity_counter"][create] = 0
    out["nextStereoFrameMinID"][create] = 0
    out["idepth"][create] = obs[create]
    out["idepth_var"][create] = F(0.5) * (F(0.5) * F(0.5))                             # VAR_RANDOM_INIT_INITIAL = 0.5f*MAX_VAR
    out["idepth_smoothed"][create] = -1
    out["idepth_var_smoothed"][create] = -1
    return out, create


def np_regularize(hyp, remove_occlusions, validity_th, depth_smoothing_factor=1.0):
    H, W = hyp.shape
    out = hyp.copy()
    valid = hyp["isValid"] > 0
    reg_dist_var = F(0.075) * F(0.075) * F(depth_smoothing_factor) * F(depth_smoothing_factor)
    pvalid = _pad(valid, 2, False)
    pid, pvar, pvc = _pad(hyp["idepth"], 2), _pad(hyp["idepth_var"], 2), _pad(hyp["validity_counter"], 2)
    cid, cvar = hyp["idepth"], hyp["idepth_var"]
    s = np.zeros((H, W), F); val_sum = np.zeros((H, W), F); s_ivar = np.zeros((H, W), F)
    n_occ = np.zeros((H, W), np.int32); n_not = np.zeros((H, W), np.int32)
    with np.errstate(all="ignore"):
        for dx in range(-2, 3):          # dx outer, dy inner (:782-783)
            for dy in range(-2, 3):
                sl = (slice(2 + dy, 2 + dy + H), slice(2 + dx, 2 + dx + W))
                sv, sid, svar = pvalid[sl], pid[sl], pvar[sl]
                diff = sid - cid
                far = (F(1.0) * F(1.0)) * diff * diff > svar + cvar
                n_occ += (sv & far & (sid > cid)).astype(np.int32)
                use = sv & ~far
                val_sum = np.where(use, val_sum + pvc[sl].astype(F), val_sum)
                n_not += use.astype(np.int32)
                dist_fac = F(dx * dx + dy * dy) * reg_dist_var
                ivar = F(1.0) / (svar + dist_fac)
                s = np.where(use, s + sid * ivar, s)
                s_ivar = np.where(use, s_ivar + ivar, s_ivar)
        yy, xx = np.mgrid[0:H, 0:W]
        centre = valid & (xx >= 2) & (xx < W - 2) & (yy >= 2) & (yy < H - 2)
        fail = centre & (val_sum < F(validity_th))
        occl = centre & ~fail & bool(remove_occlusions) & (n_occ > n_not)
        ok = centre & ~fail & ~occl
        out["isValid"][fail | occl] = 0
        out["blacklisted"][fail] -= 1
        out["idepth_smoothed"][ok] = unzero(s / s_ivar)[ok]
        out["idepth_var_smoothed"][ok] = (F(1.0) / s_ivar)[ok]
    return out


def _mapped(oracle, seq, frames, n_updates=2):
    kf = oracle.Frame(0, frames[0][0], seq.K)
    dm = oracle.DepthMap(seq.w, seq.h, seq.K)
    dm.initializeRandomly(kf)
    keep = [kf]
    for k in (4, 6)[:n_updates]:
        f = oracle.Frame(k, frames[k][0], seq.K)
        f.set_thisToParent(np.concatenate([seq.frame_to_ref_qt(k), [1.0]]), kf)
        keep.append(f)
        dm.updateKeyframe([f])
    return kf, dm, keep


def _same(a, b, fields=("isValid", "blacklisted", "validity_counter", "idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed")):
    va = a["isValid"] > 0
    assert np.array_equal(va, b["isValid"] > 0)
    assert np.array_equal(a["blacklisted"], b["blacklisted"])
    for f in fields[2:]:
        x, y = a[f][va], b[f][va]
        assert x.tobytes() == y.tobytes(), f


@pytest.mark.parametrize("occl,th", [(False, 24), (True, 24), (False, 100)])
def test_regularize_bit_exact_against_second_restatement(oracle, seq_small, frames_small, occl, th):
    kf, dm, keep = _mapped(oracle, seq_small, frames_small)
    before = dm.current().copy()
    assert (before["isValid"] > 0).sum() > 3000 and (before["blacklisted"] < 0).any()
    dm.regularize(occl, th)
    _same(np_regularize(before, occl, th), dm.current().copy())


def test_fill_holes_and_integral_against_second_restatement(oracle, seq_small, frames_small):
    kf, dm, keep = _mapped(oracle, seq_small, frames_small)
    before = dm.current().copy()
    # punch holes: drop a third of the hypotheses, blacklist some of them so that both creation thresholds are exercised
    rng = np.random.default_rng(11)
    drop = (before["isValid"] > 0) & (rng.random(before.shape) < 0.33)
    before["isValid"][drop] = 0
    before["blacklisted"][drop & (rng.random(before.shape) < 0.3)] = -3
    dm.set_current(before)
    want, created = np_fill_holes(before, kf.maxGradients(0))
    dm.regularizeFillHoles()
    assert np.array_equal(np.array(dm.integral()).reshape(before.shape), np_integral(before))
    _same(want, dm.current().copy())
    assert created.sum() > 500 and (created & (before["blacklisted"] < -1)).sum() > 5


# ---- one tracker evaluation, SE3 and Sim3 ---------------------------------------------------------------------------
def quat_R(q):
    x, y, z, w = [np.float64(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def np_point_cloud(kf, level):
    """makePointCloud: x outer, y inner; skips var <= 0 or idepth == 0 and the 1-px border"""
    K, Ki = kf.K(level)
    Ki = Ki.reshape(3, 3)
    w, h = kf.size(level)
    idepth, var, img, grad = kf.idepth(level), kf.idepthVar(level), kf.image(level), kf.gradients(level)
    xs, ys = np.meshgrid(np.arange(1, w - 1), np.arange(1, h - 1), indexing="ij")      # x-outer order
    xs, ys = xs.ravel(), ys.ravel()
    keep = ~((var[ys, xs] <= 0) | (idepth[ys, xs] == 0))
    xs, ys = xs[keep], ys[keep]
    s = F(1.0) / idepth[ys, xs]
    pos = np.stack([s * (Ki[0, 0] * xs.astype(F) + Ki[0, 2]), s * (Ki[1, 1] * ys.astype(F) + Ki[1, 2]), s * F(1)], 1).astype(F)
    return pos, grad[ys, xs, :2].astype(F), img[ys, xs].astype(F), var[ys, xs].astype(F)


def interp43(g, x, y):
    ix, iy = x.astype(np.int32), y.astype(np.int32)
    dx, dy = x - ix.astype(F), y - iy.astype(F)
    dxdy = dx * dy
    br, bl, tr, tl = g[iy + 1, ix + 1], g[iy + 1, ix], g[iy, ix + 1], g[iy, ix]
    return (dxdy[:, None] * br + (dy - dxdy)[:, None] * bl + (dx - dxdy)[:, None] * tr + (F(1) - dx - dy + dxdy)[:, None] * tl).astype(F)


def _warp(kf, frame, level, R, t):
    pos, rgrad, col, var = np_point_cloud(kf, level)
    K = frame.K(level)[0].reshape(3, 3)
    w, h = frame.size(level)
    W = ((R[None, :, 0] * pos[:, :1] + R[None, :, 1] * pos[:, 1:2]) + R[None, :, 2] * pos[:, 2:3]) + t[None, :]
    with np.errstate(all="ignore"):
        u = (W[:, 0] / W[:, 2]) * K[0, 0] + K[0, 2]
        v = (W[:, 1] / W[:, 2]) * K[1, 1] + K[1, 2]
    ok = (u > 1) & (v > 1) & (u < w - 2) & (v < h - 2)
    return pos, rgrad, col, var, W, u, v, ok, K


def seqsum(x):
    """fp32 running sum in index order, as the reference's scalar loops accumulate"""
    return np.cumsum(x.astype(F), dtype=F)[-1] if len(x) else F(0)


def test_se3_evaluation_against_second_restatement(oracle, seq_small, frames_small):
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    kf.setDepthFromGroundTruth(frames_small[0][1])
    fr = oracle.Frame(3, frames_small[3][0], seq_small.K)
    lvl = 2
    qt = seq_small.frame_to_ref_qt(3, 0)
    r2f = np.zeros(7)
    oracle.lib().lsdo_se3d_inverse(oracle._dp(qt), oracle._dp(r2f))
    r2f32 = r2f.astype(F)
    want = oracle.se3_eval(kf, fr, lvl, r2f32, 0.97, 1.5)

    q = r2f32[:4] / np.sqrt((r2f32[:4].astype(np.float64) ** 2).sum()).astype(F)
    R = quat_R(q).astype(F)
    t = r2f32[4:].astype(F)
    pos, rgrad, col, var, W, u, v, ok, K = _warp(kf, fr, lvl, R, t)
    g = interp43(fr.gradients(lvl)[..., :3], u[ok], v[ok])
    Wk, posk, vark = W[ok], pos[ok], var[ok]
    c1 = F(0.97) * col[ok] + F(1.5)
    res = c1 - g[:, 2]
    good = res * res / (F(40 * 40) + F(0.25) * (g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1])) < 1
    assert want.warpedSize == ok.sum() and want.goodCount == good.sum() and want.badCount == (~good).sum()
    usage = np.minimum(posk[:, 2] / Wk[:, 2], F(1))
    assert abs(want.pointUsage - seqsum(usage) / F(len(pos))) <= 2e-6 * want.pointUsage
    # weights (:749-790, cameraPixelNoise2 = 16, var_weight = 1, huber_d = 3)
    px, py, pz = Wk.T
    d = F(1) / posk[:, 2]
    gx, gy = K[0, 0] * g[:, 0], K[1, 1] * g[:, 1]
    g0 = (t[0] * pz - t[2] * px) / (pz * pz * d)
    g1 = (t[1] * pz - t[2] * py) / (pz * pz * d)
    drpdd = gx * g0 + gy * g1
    w_p = F(1) / (F(16) + vark * drpdd * drpdd)
    wr = np.abs(res * np.sqrt(w_p))
    wh = np.abs(np.where(wr < F(1.5), F(1), F(1.5) / wr)).astype(F)
    assert abs(want.meanWeightedRes - seqsum(wh * w_p * res * res) / F(ok.sum())) <= 1e-5 * want.meanWeightedRes
    # Jacobian rows (:1276-1291) and LGS6 (A += J J^T w, b -= J r w, finish divides by N)
    z, z2 = F(1) / pz, F(1) / (pz * pz)
    J = np.stack([z * gx, z * gy, (-px * z2) * gx + (-py * z2) * gy,
                  ((-px * py * z2) * gx).astype(np.float64) + (-(1.0 + (py * py * z2).astype(np.float64))) * gy,
                  (1.0 + (px * px * z2).astype(np.float64)) * gx + ((px * py * z2) * gy).astype(np.float64),
                  (-py * z) * gx + (px * z) * gy], 1).astype(F)
    wgt = wh * w_p
    A = np.array([[seqsum(J[:, i] * J[:, j] * wgt) for j in range(6)] for i in range(6)], F) / F(ok.sum())
    b = -np.array([seqsum(J[:, i] * (res * wgt)) for i in range(6)], F) / F(ok.sum())
    Aw, bw = np.array(want.A).reshape(6, 6), np.array(want.b)
    assert np.abs(A - Aw).max() <= 2e-5 * np.abs(Aw).max()
    assert np.abs(b - bw).max() <= 2e-5 * np.abs(bw).max() + 1e-4


def test_sim3_evaluation_against_second_restatement(oracle, seq_small, frames_small):
    kfs = {}
    for k in (0, 4):
        f = oracle.Frame(k, frames_small[k][0], seq_small.K)
        f.setDepthFromGroundTruth(frames_small[k][1])
        kfs[k] = f
    lvl = 1
    f2r = np.concatenate([seq_small.frame_to_ref_qt(4, 0), [1.03]])
    r2f = np.zeros(8)
    oracle.lib().lsdo_sim3d_inverse(oracle._dp(f2r), oracle._dp(r2f))
    want = oracle.sim3_eval(kfs[0], kfs[4], lvl, r2f, 1.0, 0.0)

    Ru = quat_R(r2f[:4])
    R = (r2f[7] * Ru).astype(F)
    t = r2f[4:7].astype(F)
    pos, rgrad, col, var, W, u, v, ok, K = _warp(kfs[0], kfs[4], lvl, R, t)
    g = interp43(kfs[4].gradients(lvl)[..., :3], u[ok], v[ok])
    # roll of the unscaled rotation about the optical axis (:451-460): shortest rotation taking R*(0,0,-1) back to (0,0,-1)
    fwd = np.array([0, 0, -1.0])
    rf = Ru @ fwd
    axis = np.cross(rf, fwd)
    ang = np.arctan2(np.linalg.norm(axis), rf @ fwd)
    if np.linalg.norm(axis) > 1e-12:
        k = axis / np.linalg.norm(axis)
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        back = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    else:
        back = np.eye(3)
    roll = (back @ Ru).astype(F)
    Wk, posk, vark, rg = W[ok], pos[ok], var[ok], rgrad[ok]
    px, py, pz = Wk.T
    gx = K[0, 0] * F(0.5) * (g[:, 0] + (roll[0, 0] * rg[:, 0] + roll[0, 1] * rg[:, 1]))
    gy = K[1, 1] * F(0.5) * (g[:, 1] + (roll[1, 0] * rg[:, 0] + roll[1, 1] * rg[:, 1]))
    rp = (F(1.0) * col[ok] + F(0.0)) - g[:, 2]
    w, h = kfs[4].size(lvl)
    idx_r = (u[ok] + F(0.5)).astype(np.int32), (v[ok] + F(0.5)).astype(np.int32)
    fvar = kfs[4].idepthVar(lvl)[idx_r[1], idx_r[0]]
    has_d = fvar > 0
    rd = np.where(has_d, F(1) / pz - kfs[4].idepth(lvl)[idx_r[1], idx_r[0]], F(-1)).astype(F)
    sv = np.where(has_d, fvar, F(-1)).astype(F)
    assert want.warpedSize == ok.sum() and want.numTermsP == ok.sum() and want.numTermsD == has_d.sum()
    d = F(1) / posk[:, 2]
    with np.errstate(all="ignore"):
        g0 = (t[0] * pz - t[2] * px) / (pz * pz * d)
        g1 = (t[1] * pz - t[2] * py) / (pz * pz * d)
        g2 = (pz - t[2]) / (pz * pz * d)
        drpdd = gx * g0 + gy * g1
        w_p = F(1) / (F(16) + vark * drpdd * drpdd)
        w_d = F(1) / (sv + g2 * g2 * vark)
        wrd, wrp = np.abs(rd * np.sqrt(w_d)), np.abs(rp * np.sqrt(w_p))
        tot = np.where(has_d, wrd + wrp, wrp)
        wh = np.abs(np.where(tot < F(3), F(1), F(3) / tot)).astype(F)
    sumD, sumP = seqsum((wh * w_d * rd * rd)[has_d]), seqsum(wh * w_p * rp * rp)
    assert abs(want.sumResD - sumD) <= 2e-5 * want.sumResD and abs(want.sumResP - sumP) <= 2e-5 * want.sumResP
    assert abs(want.mean - (sumD + sumP) / F(has_d.sum() + ok.sum())) <= 2e-5 * want.mean
    wp, wd = wh * w_p, np.where(has_d, wh * w_d, F(0)).astype(F)
    z, z2 = F(1) / pz, F(1) / (pz * pz)
    v6 = np.stack([z * gx, z * gy, (-px * z2) * gx + (-py * z2) * gy,
                   ((-px * py * z2) * gx).astype(np.float64) + (-(1.0 + (py * py * z2).astype(np.float64))) * gy,
                   (1.0 + (px * px * z2).astype(np.float64)) * gx + ((px * py * z2) * gy).astype(np.float64),
                   (-py * z) * gx + (px * z) * gy], 1).astype(F)
    v4 = np.stack([z2, z2 * py, -z2 * px, z], 1).astype(F)
    A7 = np.zeros((7, 7), F); b7 = np.zeros(7, F)
    A7[:6, :6] = [[seqsum(v6[:, i] * v6[:, j] * wp) for j in range(6)] for i in range(6)]
    b7[:6] = [-seqsum(v6[:, i] * (rp * wp)) for i in range(6)]
    remap = [2, 3, 4, 6]
    for i in range(4):
        b7[remap[i]] += -seqsum(v4[:, i] * (rd * wd))
        for j in range(4):
            A7[remap[i], remap[j]] += seqsum(v4[:, i] * v4[:, j] * wd)
    Aw, bw = np.array(want.A).reshape(7, 7), np.array(want.b)
    assert np.abs(A7 - Aw).max() <= 5e-5 * np.abs(Aw).max()
    assert np.abs(b7 - bw).max() <= 5e-5 * np.abs(bw).max() + 1e-3 * np.sqrt(np.abs(Aw).max())
    assert want.num_constraints == 2 * ok.sum()


# ---- per-pixel stereo: observeDepthCreate / observeDepthUpdate / doLineStereo ------------------------------------------
def _hyp_dict(h):
    return {k: (h[k].item() if k in ("isValid", "blacklisted", "validity_counter") else F(h[k])) for k in h.dtype.names}


def _cam(kf):
    K, Ki = kf.K(0)
    K, Ki = K.reshape(3, 3), Ki.reshape(3, 3)
    return dict(fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2], fxi=Ki[0, 0], fyi=Ki[1, 1], cxi=Ki[0, 2], cyi=Ki[1, 2])


@pytest.mark.parametrize("init", ["empty", "gt", "random"])
def test_observe_depth_pixels_against_second_restatement(oracle, seq_small, frames_small, init):
    """observeDepthRow on sampled pixels, pure-Python fp32 restatement vs the C oracle: every field bit for bit"""
    from tests import restate_stereo as rs
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    dm = oracle.DepthMap(seq_small.w, seq_small.h, seq_small.K)
    if init == "gt":
        kf.setDepthFromGroundTruth(frames_small[0][1])
        dm.initializeFromGTDepth(kf)
        dm.regularize(False, 24)                    # gives idepth_smoothed / var_smoothed their filtered values
    else:
        dm.initializeRandomly(kf)
        if init == "empty":
            cur = dm.current().copy()
            cur["isValid"] = 0
            dm.set_current(cur)
    k = 5
    fr = oracle.Frame(k, frames_small[k][0], seq_small.K)
    qts = np.concatenate([seq_small.frame_to_ref_qt(k), [1.0]])
    fr.set_thisToParent(qts, kf)
    before = dm.current().copy()
    dm.observeDepth([fr])
    after = dm.current().copy()

    G = dict(minUseGrad=F(5.0), cameraPixelNoise2=F(16.0), useSubpixelStereo=True, allowNegativeIdepths=True)
    ref = rs.Ref(seq_small.K, qts, fr.image(0), k, initial_tracked_residual=0.0, good_mask=None)
    refs = dict(oldest=ref, newest=ref, by_id=[ref], offset=k)
    state = dict(reactivated=False, numTracked=0, numMapped=0)
    cam = _cam(kf)
    img, grad, mg = kf.image(0), kf.gradients(0), kf.maxGradients(0)
    rng = np.random.default_rng(5)
    ys, xs = np.nonzero(mg[3:-3, 3:-3] >= 5.0)
    pick = rng.choice(len(xs), 350, replace=False)
    changed = 0
    for x, y in zip(xs[pick] + 3, ys[pick] + 3):
        got = rs.observe_pixel(cam, G, img, grad, mg, _hyp_dict(before[y, x]), int(x), int(y), refs, state)
        want = after[y, x]
        assert int(got["isValid"]) == int(want["isValid"]) and int(got["blacklisted"]) == int(want["blacklisted"]), (x, y)
        if want["isValid"]:
            assert int(got["validity_counter"]) == int(want["validity_counter"]), (x, y)
            for f in ("idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed", "nextStereoFrameMinID"):
                assert F(got[f]).tobytes() == F(want[f]).tobytes(), (x, y, f, got[f], want[f])
        changed += int(before[y, x].tobytes() != want.tobytes())
    assert changed > 60


@pytest.mark.parametrize("with_mask", [False, True])
def test_propagate_depth_against_second_restatement(oracle, seq_small, frames_small, with_mask):
    """DepthMap::propagateDepth (order-dependent raster scan with merge / occlusion handling), both admission tests:
    the tracker's refPixelWasGood mask and the colour test for untracked frames"""
    from tests import restate_stereo as rs
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    kf.setDepthFromGroundTruth(frames_small[0][1])
    dm = oracle.DepthMap(seq_small.w, seq_small.h, seq_small.K)
    dm.initializeFromGTDepth(kf)
    dm.regularize(False, 24)
    k = 9
    nf = oracle.Frame(k, frames_small[k][0], seq_small.K)
    if with_mask:
        r = oracle.se3_track(kf, nf, np.array([0, 0, 0, 1, 0, 0, 0], np.float64))
        assert not r.diverged
        qts = nf.thisToParent()
        mask = nf.refPixelWasGood().copy().astype(bool)
        assert (~mask).sum() > 50                       # the mask really rejects some pixels
    else:
        qts = np.concatenate([seq_small.frame_to_ref_qt(k), [1.0]])
        nf.set_thisToParent(qts, kf)
        mask = None
    before = dm.current().copy()
    dm.propagateDepth(nf)
    after = dm.current().copy()
    want = rs.propagate_depth(_cam(kf), kf.image(0), nf.image(0), nf.maxGradients(0), before, qts, mask)
    va = after["isValid"] > 0
    assert np.array_equal(va, want["isValid"] > 0)
    assert va.sum() > 5000
    assert np.array_equal(after["blacklisted"], want["blacklisted"])
    for f in ("validity_counter", "idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed", "nextStereoFrameMinID"):
        assert after[f][va].tobytes() == want[f][va].tobytes(), f


# ---- Frame builders ------------------------------------------------------------------------------------------------
def np_max_gradients(grad, min_create=F(5.0)):
    """Frame::buildMaxGradients, DataStructures/Frame.cpp:690-767: three LINEAR-index sweeps (x wraps across rows);
    never-written cells are zero (pool buffers defined zero-filled)"""
    h, w = grad.shape[:2]
    n = w * h
    g = grad.reshape(n, 4)
    a = np.zeros(n, F)
    a[w:n - w] = np.sqrt(g[w:n - w, 0] * g[w:n - w, 0] + g[w:n - w, 1] * g[w:n - w, 1])
    t = np.zeros(n, F)
    i = np.arange(w + 1, n - w - 1)
    t[i] = np.maximum(np.maximum(a[i - w], a[i]), a[i + w])
    out = a.copy()
    out[i] = np.maximum(np.maximum(t[i - 1], t[i]), t[i + 1])
    return out.reshape(h, w), int((out[i] >= min_create).sum())


def np_idepth_level(idepth, var):
    """Frame::buildIDepthAndIDepthVar for one level, DataStructures/Frame.cpp:775-877 (source order: tl, tr, bl, br)"""
    h, w = idepth.shape[0] // 2, idepth.shape[1] // 2
    isum = np.zeros((h, w), F); dsum = np.zeros((h, w), F); num = np.zeros((h, w), np.int32)
    with np.errstate(all="ignore"):
        for oy, ox in ((0, 0), (0, 1), (1, 0), (1, 1)):
            v, d = var[oy::2, ox::2][:h, :w], idepth[oy::2, ox::2][:h, :w]
            m = v > 0
            ivar = F(1.0) / v
            isum = np.where(m, isum + ivar, isum)
            dsum = np.where(m, dsum + ivar * d, dsum)
            num += m
        depth = isum / dsum
        oid = np.where(num > 0, F(1.0) / depth, F(-1)).astype(F)
        ovar = np.where(num > 0, num.astype(F) / isum, F(-1)).astype(F)
    return oid, ovar


def test_frame_builders_against_second_restatement(oracle, seq_small, frames_small):
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    mg, mappable = np_max_gradients(kf.gradients(0))
    assert kf.maxGradients(0).tobytes() == mg.tobytes()
    assert oracle.lib().lsdo_frame_numMappablePixels(kf.ptr) == mappable
    dm = oracle.DepthMap(seq_small.w, seq_small.h, seq_small.K)
    dm.initializeRandomly(kf)                       # ends with Frame::setDepth(currentDepthMap)
    cur = dm.current().copy()
    ok = (cur["isValid"] > 0) & (cur["idepth_smoothed"].astype(np.float64) >= -0.05)
    assert np.array_equal(kf.idepth(0), np.where(ok, cur["idepth_smoothed"], F(-1)))
    assert np.array_equal(kf.idepthVar(0), np.where(ok, cur["idepth_var_smoothed"], F(-1)))
    assert oracle.lib().lsdo_frame_numPoints(kf.ptr) == ok.sum()
    mean = seqsum(cur["idepth_smoothed"][ok]) / F(ok.sum())               # row-major fp32 running sum, Frame.cpp:217-234
    assert abs(oracle.lib().lsdo_frame_meanIdepth(kf.ptr) - mean) <= 1e-6 * mean
    idl, vl = kf.idepth(0).copy(), kf.idepthVar(0).copy()
    for lvl in range(1, 5):
        idl, vl = np_idepth_level(idl, vl)
        assert kf.idepth(lvl).tobytes() == idl.tobytes(), lvl
        assert kf.idepthVar(lvl).tobytes() == vl.tobytes(), lvl


# ---- LM control flow of the two trackers (second restatement of the LOOP; evaluations and small solves are the oracle's) ----
def _se3f(oracle, name, *arrs):
    out = np.zeros(7, F)
    getattr(oracle.lib(), name)(*[oracle._fp(np.ascontiguousarray(a, F)) for a in arrs], oracle._fp(out))
    return out


def py_track_frame(oracle, kf, fr, init_f2r, settings, use_affine=True, levels=(4, 3, 2, 1), init_is_ref_to_frame=False):
    """SE3Tracker::trackFrame, Tracking/SE3Tracker.cpp:280-486: level schedule, LM damping, accept / reject / convergence.
    levels=(4,), init_is_ref_to_frame=True: the loop of trackFrameOnPermaref (:162-272), which returns referenceToFrame."""
    L = oracle.lib()
    inv = np.zeros(7)
    if init_is_ref_to_frame:
        inv[:] = init_f2r
    else:
        L.lsdo_se3d_inverse(oracle._dp(np.ascontiguousarray(init_f2r, np.float64)), oracle._dp(inv))
    r2f = inv.astype(F)
    r2f[:4] = r2f[:4] / np.sqrt(F(F(F(r2f[0] * r2f[0]) + F(r2f[1] * r2f[1])) + F(r2f[2] * r2f[2])) + F(r2f[3] * r2f[3]))   # cast<float>() normalises
    a, b = F(1), F(0)
    n_res, n_upd = [0] * 5, [0] * 5
    last_residual = F(0)
    W, H = kf.w, kf.h
    for lvl in levels:
        ev = oracle.se3_eval(kf, fr, lvl, r2f, a, b, settings, lvl == 1)
        if ev.warpedSize < F(0.01) * (W >> lvl) * (H >> lvl):
            return None, n_res, n_upd
        if use_affine:
            a, b = F(ev.affine_a_lastIt), F(ev.affine_b_lastIt)
        last_err = F(ev.meanWeightedRes)
        ls = ev
        n_res[lvl] += 1
        lam = F(settings.lambdaInitial[lvl])
        it = 0
        while it < settings.maxItsPerLvl[lvl]:
            n_upd[lvl] += 1
            inc_try = 0
            while True:
                A = np.array(ls.A, F).reshape(6, 6).copy()
                for i in range(6):
                    A[i, i] = F(A[i, i] * F(1 + lam))
                rhs = (-np.array(ls.b, F)).astype(F)
                inc = np.zeros(6, F)
                L.lsdo_ldlt6_solve(oracle._fp(np.ascontiguousarray(A.reshape(36))), oracle._fp(rhs), oracle._fp(inc))
                inc_try += 1
                new = _se3f(oracle, "lsdo_se3f_mul", _se3f(oracle, "lsdo_se3f_exp", inc), r2f)
                ev = oracle.se3_eval(kf, fr, lvl, new, a, b, settings, lvl == 1)
                if ev.warpedSize < F(0.01) * (W >> lvl) * (H >> lvl):
                    return None, n_res, n_upd
                err = F(ev.meanWeightedRes)
                n_res[lvl] += 1
                if err < last_err:
                    r2f = new
                    if use_affine:
                        a, b = F(ev.affine_a_lastIt), F(ev.affine_b_lastIt)
                    if F(err / last_err) > F(settings.convergenceEps[lvl]):
                        it = settings.maxItsPerLvl[lvl]
                    last_residual = last_err = err
                    ls = ev
                    lam = F(0) if np.float64(lam) <= 0.2 else F(lam * F(settings.lambdaSuccessFac))   # float compared with the double literal 0.2
                    break
                dot = F(0)
                for i in range(6):
                    dot = F(dot + F(inc[i] * inc[i]))
                if not (dot > F(settings.stepSizeMin[lvl])):
                    it = settings.maxItsPerLvl[lvl]
                    break
                lam = F(0.2) if lam == 0 else F(np.float64(lam) * np.float64(F(settings.lambdaFailFac)) ** inc_try)
            it += 1
    if init_is_ref_to_frame:
        return r2f, n_res, n_upd
    f2r = _se3f(oracle, "lsdo_se3f_inverse", r2f)
    return f2r, n_res, n_upd


@pytest.mark.parametrize("k", [2, 7])
def test_track_frame_control_flow_against_second_restatement(oracle, seq_small, frames_small, k):
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    kf.setDepthFromGroundTruth(frames_small[0][1])
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    s = oracle.default_track_settings()
    fa, fb = oracle.Frame(k, frames_small[k][0], seq_small.K), oracle.Frame(k, frames_small[k][0], seq_small.K)
    want = oracle.se3_track(kf, fa, ident, s)
    got, n_res, n_upd = py_track_frame(oracle, kf, fb, ident, s)
    assert n_res == list(want.numCalcResidualCalls) and n_upd == list(want.numCalcWarpUpdateCalls)
    assert sum(n_res) > sum(n_upd) + 4 > 8                      # the schedule really contains rejected tries
    q = np.array(want.frameToRef_qt)
    assert np.abs(got.astype(np.float64)[4:] - q[4:]).max() == 0 and np.abs(got.astype(np.float64)[:4] - q[:4]).max() < 1e-7
    assert np.array_equal(fa.refPixelWasGood(), fb.refPixelWasGood())


def test_track_frame_sim3_control_flow_against_second_restatement(oracle, seq_small, frames_small):
    """Sim3Tracker::trackFrameSim3, Tracking/Sim3Tracker.cpp:149-382, driven from Python over lsdo_sim3_eval"""
    L = oracle.lib()
    kfs = {}
    for k in (0, 4):
        f = oracle.Frame(k, frames_small[k][0], seq_small.K)
        f.setDepthFromGroundTruth(frames_small[k][1])
        kfs[k] = f
    init = np.concatenate([seq_small.frame_to_ref_qt(4, 0), [1.02]])
    init[4:7] += [0.01, -0.005, 0.004]
    s = oracle.default_track_settings(main_tracker=False)
    want = oracle.sim3_track(kfs[0], kfs[4], init, 4, 1, s)

    def qts(fn, *a):
        out = np.zeros(8)
        getattr(L, fn)(*[oracle._dp(np.ascontiguousarray(x, np.float64)) for x in a], oracle._dp(out))
        return out
    r2f = qts("lsdo_sim3d_inverse", init)
    a, b = 1.0, 0.0
    n_res, n_upd = [0] * 5, [0] * 5
    W, H = kfs[0].w, kfs[0].h
    up_to_date, final = False, None
    for lvl in range(4, 0, -1):
        if s.maxItsPerLvl[lvl] == 0:
            continue
        ev = oracle.sim3_eval(kfs[0], kfs[4], lvl, r2f, a, b, s)
        assert not (ev.warpedSize < 0.5 * F(0.01) * (W >> lvl) * (H >> lvl) or ev.warpedSize < 10)
        last = ev
        n_res[lvl] += 1
        a, b = ev.affine_a_lastIt, ev.affine_b_lastIt
        lam = F(s.lambdaInitial[lvl])
        up_to_date = False
        ls = ev
        it = 0
        while it < s.maxItsPerLvl[lvl]:
            up_to_date = True
            n_upd[lvl] += 1
            inc_try = 0
            while True:
                nc = F(ls.num_constraints)
                A = (np.array(ls.A, F) / nc).astype(F).reshape(7, 7)
                for i in range(7):
                    A[i, i] = F(A[i, i] * F(1 + lam))
                rhs = (-np.array(ls.b, F) / nc).astype(F)
                inc = np.zeros(7, F)
                L.lsdo_ldlt7_solve(oracle._fp(np.ascontiguousarray(A.reshape(49))), oracle._fp(rhs), oracle._fp(inc))
                inc_try += 1
                abs_inc = F(0)
                for i in range(7):
                    abs_inc = F(abs_inc + F(inc[i] * inc[i]))
                assert abs_inc >= 0 and abs_inc < 1
                new = qts("lsdo_sim3d_mul", qts("lsdo_sim3d_exp", inc.astype(np.float64)), r2f)
                ev = oracle.sim3_eval(kfs[0], kfs[4], lvl, new, a, b, s)
                n_res[lvl] += 1
                if ev.mean < last.mean:
                    r2f = new
                    up_to_date = False
                    a, b = ev.affine_a_lastIt, ev.affine_b_lastIt
                    if F(F(ev.mean) / F(last.mean)) > F(s.convergenceEps[lvl]):
                        it = s.maxItsPerLvl[lvl]
                    final = last = ev
                    ls = ev
                    lam = F(0) if np.float64(lam) <= 0.2 else F(lam * F(s.lambdaSuccessFac))
                    break
                if not (abs_inc > F(s.stepSizeMin[lvl])):
                    it = s.maxItsPerLvl[lvl]
                    break
                lam = F(0.2) if lam == 0 else F(np.float64(lam) * np.float64(F(s.lambdaFailFac)) ** inc_try)
            it += 1
    if not up_to_date:
        final = ls = oracle.sim3_eval(kfs[0], kfs[4], 1, r2f, a, b, s)
    assert n_res == list(want.numCalcResidualCalls) and n_upd == list(want.numCalcWarpUpdateCalls)
    got = qts("lsdo_sim3d_inverse", r2f)
    assert np.abs(got - np.array(want.frameToRef_qts)).max() < 1e-6
    assert abs(final.mean - want.lastResidual) <= 1e-5 * want.lastResidual
    assert np.abs(np.array(ls.A) - np.array(want.lastSim3Hessian)).max() <= 1e-4 * np.abs(np.array(want.lastSim3Hessian)).max()


def test_create_keyframe_against_second_restatement(oracle, seq_small, frames_small):
    """DepthMap::createKeyFrame, DepthEstimation/DepthMap.cpp:1222-1327, composed from the second restatements:
    propagateDepth -> regularize<true>(KEEP) -> fillHoles -> regularize<false>(KEEP) -> rescale to mean idepth 1 -> setDepth"""
    from tests import restate_stereo as rs
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    kf.setDepthFromGroundTruth(frames_small[0][1])
    dm = oracle.DepthMap(seq_small.w, seq_small.h, seq_small.K)
    dm.initializeFromGTDepth(kf)
    dm.regularize(False, 24)
    k = 8
    nf = oracle.Frame(k, frames_small[k][0], seq_small.K)
    r = oracle.se3_track(kf, nf, np.array([0, 0, 0, 1, 0, 0, 0], np.float64))
    qts = nf.thisToParent().copy()
    mask = nf.refPixelWasGood().copy().astype(bool)
    before = dm.current().copy()
    dm.createKeyFrame(nf)
    after = dm.current().copy()

    cur = rs.propagate_depth(_cam(kf), kf.image(0), nf.image(0), nf.maxGradients(0), before, qts, mask)
    cur = np_regularize(cur, True, 24)
    cur, _ = np_fill_holes(cur, nf.maxGradients(0))
    cur = np_regularize(cur, False, 24)
    va = cur["isValid"] > 0
    s = seqsum(cur["idepth_smoothed"][va])                    # row-major fp32 running sum, :1286-1293
    f = F(F(va.sum()) / s)
    f2 = F(f * f)
    for name, fac in (("idepth", f), ("idepth_smoothed", f), ("idepth_var", f2), ("idepth_var_smoothed", f2)):
        cur[name][va] = (cur[name][va] * fac).astype(F)
    assert np.array_equal(va, after["isValid"] > 0) and va.sum() > 5000
    assert np.array_equal(cur["blacklisted"], after["blacklisted"])
    for name in ("validity_counter", "idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed"):
        assert cur[name][va].tobytes() == after[name][va].tobytes(), name
    new_pose = nf.thisToParent()
    assert abs(new_pose[7] - np.float64(f)) == 0                 # sim3FromSE3(oldToNew.inverse(), rescaleFactor), :1305
    assert np.abs(new_pose[4:7] - qts[4:7]).max() < 1e-12 and abs(abs(np.dot(new_pose[:4], qts[:4])) - 1) < 1e-12
    ok = va & (cur["idepth_smoothed"].astype(np.float64) >= -0.05)
    assert np.array_equal(nf.idepth(0), np.where(ok, cur["idepth_smoothed"], F(-1)))


def test_perma_ref_overlap_against_second_restatement(oracle, seq_small, frames_small):
    """SE3Tracker::checkPermaRefOverlap, Tracking/SE3Tracker.cpp:121-157 over the Frame::setPermaRef snapshot (level 4)"""
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    kf.setDepthFromGroundTruth(frames_small[0][1])
    pr = oracle.PermaRef(kf)
    pos, _, col, var = np_point_cloud(kf, 4)
    assert pr.n == len(pos) and np.array_equal(pr.pos[:pr.n], pos) and np.array_equal(pr.colvar[:pr.n, 0], col) and np.array_equal(pr.colvar[:pr.n, 1], var)
    K4 = kf.K(4)[0].reshape(3, 3)
    w2, h2 = kf.size(4)[0] - 1, kf.size(4)[1] - 1
    rng = np.random.default_rng(2)
    for _ in range(6):
        a = rng.normal(0, [0.2, 0.2, 0.3, 0.05, 0.05, 0.05])
        qt = np.zeros(7)
        oracle.lib().lsdo_se3d_exp(oracle._dp(a), oracle._dp(qt))
        q32 = qt.astype(F)
        q32[:4] = q32[:4] / np.sqrt(F(F(F(q32[0] * q32[0]) + F(q32[1] * q32[1])) + F(q32[2] * q32[2])) + F(q32[3] * q32[3]))
        R, t = quat_R(q32[:4]).astype(F), q32[4:]
        # the float rotation matrix of a float quaternion: redo in fp32 to match Eigen's toRotationMatrix on SE3f
        x, y, z, w = q32[:4]
        tx, ty, tz = F(2) * x, F(2) * y, F(2) * z
        R = np.array([[F(1) - (ty * y + tz * z), ty * x - tz * w, tz * x + ty * w],
                      [ty * x + tz * w, F(1) - (tx * x + tz * z), tz * y - tx * w],
                      [tz * x - ty * w, tz * y + tx * w, F(1) - (tx * x + ty * y)]], F)
        W = ((R[None, :, 0] * pos[:, :1] + R[None, :, 1] * pos[:, 1:2]) + R[None, :, 2] * pos[:, 2:3]) + t[None, :]
        with np.errstate(all="ignore"):
            u = (W[:, 0] / W[:, 2]) * K4[0, 0] + K4[0, 2]
            v = (W[:, 1] / W[:, 2]) * K4[1, 1] + K4[1, 2]
            inside = (u > 0) & (v > 0) & (u < w2) & (v < h2)
            usage = seqsum(np.minimum(pos[inside, 2] / W[inside, 2], F(1))) / F(len(pos))
        assert abs(pr.overlap(qt) - usage) <= 2e-6 * max(usage, 1e-3)


def test_track_frame_on_permaref_control_flow_against_second_restatement(oracle, seq_small, frames_small):
    """SE3Tracker::trackFrameOnPermaref, Tracking/SE3Tracker.cpp:162-272: the level-4 loop with the TestTrack settings
    (util/settings.h:379-382), re-driven over single evaluations of the same level-4 cloud"""
    kf = oracle.Frame(0, frames_small[0][0], seq_small.K)
    kf.setDepthFromGroundTruth(frames_small[0][1])
    pr = oracle.PermaRef(kf)
    fr = oracle.Frame(6, frames_small[6][0], seq_small.K)
    s = oracle.default_track_settings(main_tracker=False)
    s.lambdaInitial[4], s.stepSizeMin[4], s.convergenceEps[4], s.maxItsPerLvl[4] = 0, 1e-3, 0.98, 5
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)
    want = pr.track(fr, ident)
    got, n_res, n_upd = py_track_frame(oracle, kf, fr, ident, s, levels=(4,), init_is_ref_to_frame=True)
    assert not want.diverged
    assert n_res[4] == want.numCalcResidualCalls[4] and n_upd[4] == want.numCalcWarpUpdateCalls[4] and n_res[4] >= 3
    q = np.array(want.frameToRef_qt)                  # holds referenceToFrame for this call (:271)
    assert np.abs(got.astype(np.float64)[4:] - q[4:]).max() == 0 and np.abs(got.astype(np.float64)[:4] - q[:4]).max() < 1e-7


def test_initialisers_against_second_restatement(oracle, seq_small, frames_small):
    """Frame::setDepthFromGroundTruth (Frame.cpp:245-293), DepthMap::initializeFromGTDepth (DepthMap.cpp:965-1018) and
    initializeRandomly (:883-918, glibc rand() re-seeded to its default state for both sides)"""
    import ctypes
    img, depth = frames_small[0]
    kf = oracle.Frame(0, img, seq_small.K)
    mg = kf.maxGradients(0)
    H, W = mg.shape
    depth = depth.copy()
    depth[50:60, 70:90] = np.nan
    depth[100:105, 10:40] = -1.0
    kf.setDepthFromGroundTruth(depth, 2.0)
    yy, xx = np.mgrid[0:H, 0:W]
    with np.errstate(all="ignore"):
        ok = (xx > 0) & (xx < W - 1) & (yy > 0) & (yy < H - 1) & (mg >= F(5.0)) & ~np.isnan(depth) & (depth > 0)
        want_id = np.where(ok, F(1.0) / depth, F(-1)).astype(F)
    want_var = np.where(ok, F(F(0.01) * F(0.01)) * F(2.0), F(-1)).astype(F)       # VAR_GT_INIT_INITIAL * cov_scale
    assert kf.idepth(0).tobytes() == want_id.tobytes() and kf.idepthVar(0).tobytes() == want_var.tobytes()

    dm = oracle.DepthMap(W, H, seq_small.K)
    dm.initializeFromGTDepth(kf)
    cur = dm.current().copy()
    v = ~np.isnan(want_id) & (want_id > 0)
    assert np.array_equal(cur["isValid"] > 0, v)
    assert np.array_equal(cur["idepth"][v], want_id[v]) and np.array_equal(cur["idepth_smoothed"][v], want_id[v])
    assert (cur["idepth_var"][v] == F(0.01) * F(0.01)).all() and (cur["validity_counter"][v] == 20).all() and not cur["blacklisted"].any()

    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    kf2 = oracle.Frame(1, img, seq_small.K)
    dm2 = oracle.DepthMap(W, H, seq_small.K)
    dm2.initializeRandomly(kf2)
    got = dm2.current().copy()
    libc.srand(1)
    sel = (mg > F(5.0))[1:H - 1, 1:W - 1]
    n = int(sel.sum())
    r = np.array([libc.rand() for _ in range(n)], np.int64)                   # row-major order of the reference's loops
    idepth = (F(0.5) + F(1.0) * ((r % 100001).astype(F) / F(100000.0))).astype(F)
    inner = got[1:H - 1, 1:W - 1]
    assert np.array_equal(inner["isValid"] > 0, sel)
    assert inner["idepth"][sel].tobytes() == idepth.tobytes() and inner["idepth_smoothed"][sel].tobytes() == idepth.tobytes()
    assert (inner["idepth_var"][sel] == F(0.5) * (F(0.5) * F(0.5))).all() and (inner["validity_counter"][sel] == 20).all()
