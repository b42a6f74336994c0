"""Second restatement (see tests/test_oracle_independent.py) of the per-pixel stereo update of DepthMap, written from the
reference text with numpy float32 SCALARS (every operation rounds to fp32 like the reference's `float` arithmetic; where the
reference mixes in double literals the promotion is made explicit):

    observeDepthRow            DepthEstimation/DepthMap.cpp:108-139
    makeAndCheckEPL            :176-231
    observeDepthCreate         :234-291
    observeDepthUpdate         :294-473
    doLineStereo               :1442-1972   (Release semantics: enablePrintDebugInfo == false)
    Frame::prepareForStereoWith  DataStructures/Frame.cpp:295-317
    getInterpolatedElement / 42  util/globalFuncs.h:43-61, 95-109

Pure Python, one pixel at a time: slow, used on a few hundred sampled pixels only.  Conventions where Eigen versions differ
are the ones stated in oracle/lsd_oracle.h (matrix * vector summed left to right, dot() as a0b0 + (a1b1 + a2b2), vector /
scalar as a true division)."""
import numpy as np

F = np.float32
D = np.float64
NAN, INF = F(np.nan), F(np.inf)

MIN_DEPTH = F(0.05)
MAX_EPL_LENGTH_CROP, MIN_EPL_LENGTH_CROP = F(30.0), F(3.0)
GRADIENT_SAMPLE_DIST = F(1.0)
SAMPLE_POINT_TO_BORDER = 7
MAX_ERROR_STEREO, MIN_DISTANCE_ERROR_STEREO = F(1300.0), F(1.5)
STEREO_EPL_VAR_FAC = F(2.0)
DIVISION_EPS = F(1e-10)
MAX_VAR = F(0.5) * F(0.5)
SUCC_VAR_INC_FAC, FAIL_VAR_INC_FAC = F(1.01), F(1.1)
MIN_BLACKLIST = -1


def unzero(v):
    v = D(v)
    return F((D(-1e-10) if v > -1e-10 else v) if v < 0 else (D(1e-10) if v < 1e-10 else v))


def interp(img, x, y):
    """img: 2-D float32 array addressed [row, col]; the reference indexes mat + ix + iy*width"""
    flat = img.reshape(-1)
    w = img.shape[1]
    ix, iy = int(x), int(y)
    dx, dy = F(x - F(ix)), F(y - F(iy))
    dxdy = F(dx * dy)
    b = ix + iy * w
    return F(F(F(dxdy * flat[b + 1 + w] + F(F(dy - dxdy) * flat[b + w])) + F(F(dx - dxdy) * flat[b + 1])) + F(F(F(F(F(1) - dx) - dy) + dxdy) * flat[b]))


def interp42(grad4, x, y):
    w = grad4.shape[1]
    flat = grad4.reshape(-1, 4)
    ix, iy = int(x), int(y)
    dx, dy = F(x - F(ix)), F(y - F(iy))
    dxdy = F(dx * dy)
    b = ix + iy * w
    out = []
    for c in (0, 1):
        out.append(F(F(F(dxdy * flat[b + 1 + w, c] + F(F(dy - dxdy) * flat[b + w, c])) + F(F(dx - dxdy) * flat[b + 1, c]))
                     + F(F(F(F(F(1) - dx) - dy) + dxdy) * flat[b, c])))
    return out


def quat_R64(q):
    x, y, z, w = [D(v) for v in q]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx], [txz - twy, tyz + twx, 1 - (txx + tyy)]], D)


class Ref:
    """what observeDepth reads from a reference frame after prepareForStereoWith(thisToOther = refToKf.inverse())"""

    def __init__(self, K, frame_to_kf_qts, image, fid, initial_tracked_residual=0.0, good_mask=None):
        # DepthMap.cpp:1094-1100: frame->prepareForStereoWith(activeKeyFrame, refToKf = frame->thisToParent_raw, K, 0):
        # this = the reference frame, other = the keyframe, thisToOther = frame -> keyframe
        q, t, s = np.asarray(frame_to_kf_qts[:4], D), np.asarray(frame_to_kf_qts[4:7], D), D(frame_to_kf_qts[7])
        R_t2o = quat_R64(q)
        # otherToThis = thisToOther.inverse(): rotation R^T, scale 1/s, translation -(1/s) R^T t   (sim3.hpp:169-173, in double)
        si = D(1) / s
        R_o2t = R_t2o.T
        t_o2t = -(si * (R_o2t @ t))
        Kf = np.asarray(K, F).reshape(3, 3)
        Rf = R_o2t.astype(F)
        KR = np.zeros((3, 3), F)
        for i in range(3):
            for j in range(3):
                KR[i, j] = F(F(F(F(Kf[i, 0] * Rf[0, j]) + F(Kf[i, 1] * Rf[1, j])) + F(Kf[i, 2] * Rf[2, j])) * F(si))
        self.K_otherToThis_R = KR
        self.otherToThis_t = np.array([F(v) for v in t_o2t], F)
        self.K_otherToThis_t = np.array([F(F(F(Kf[i, 0] * self.otherToThis_t[0]) + F(Kf[i, 1] * self.otherToThis_t[1])) + F(Kf[i, 2] * self.otherToThis_t[2]))
                                         for i in range(3)], F)
        self.thisToOther_t = np.array([F(v) for v in t], F)
        tR = (R_t2o.astype(F) * F(s)).astype(F)          # thisToOther_R = rotationMatrix().cast<float>() * scale
        self.row0, self.row1, self.row2 = tR[:, 0].copy(), tR[:, 1].copy(), tR[:, 2].copy()     # "rows" are columns (:310-312)
        self.image = image
        self.id = fid
        self.initialTrackedResidual = F(initial_tracked_residual)
        self.good_mask = good_mask


def make_and_check_epl(cam, kf_image, x, y, ref):
    fx, fy, cx, cy = cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    w = kf_image.shape[1]
    flat = kf_image.reshape(-1)
    idx = x + y * w
    epx = F(F(-fx * ref.thisToOther_t[0]) + F(ref.thisToOther_t[2] * F(F(x) - cx)))
    epy = F(F(-fy * ref.thisToOther_t[1]) + F(ref.thisToOther_t[2] * F(F(y) - cy)))
    if np.isnan(F(epx + epy)):
        return None
    l2 = F(F(epx * epx) + F(epy * epy))
    if l2 < F(1.0) * F(1.0):
        return None
    gx = F(flat[idx + 1] - flat[idx - 1])
    gy = F(flat[idx + w] - flat[idx - w])
    e = F(F(gx * epx) + F(gy * epy))
    e = F(F(e * e) / l2)
    if e < F(2.0) * F(2.0):
        return None
    with np.errstate(all="ignore"):
        if F(e / F(F(gx * gx) + F(gy * gy))) < F(0.3) * F(0.3):
            return None
    fac = F(GRADIENT_SAMPLE_DIST / np.sqrt(l2))
    return F(epx * fac), F(epy * fac)


def dot3(a, b):
    return F(F(a[0] * b[0]) + F(F(a[1] * b[1]) + F(a[2] * b[2])))


def do_line_stereo(cam, G, kf_image, kf_grad, u, v, epxn, epyn, min_idepth, prior_idepth, max_idepth, ref):
    """-> (error, result_idepth, result_var, result_eplLength); the results are None on the early returns"""
    width, height = kf_image.shape[1], kf_image.shape[0]
    fxi, fyi, cxi, cyi = cam["fxi"], cam["fyi"], cam["cxi"], cam["cyi"]
    KR, Kt = ref.K_otherToThis_R, ref.K_otherToThis_t
    with np.errstate(all="ignore"):
        KinvP = [F(F(fxi * u) + cxi), F(F(fyi * v) + cyi), F(1.0)]
        pInf = [F(F(F(KR[i, 0] * KinvP[0]) + F(KR[i, 1] * KinvP[1])) + F(KR[i, 2] * KinvP[2])) for i in range(3)]
        pReal = [F(F(pInf[i] / prior_idepth) + Kt[i]) for i in range(3)]
        rescale = F(pReal[2] * prior_idepth)
        firstX, firstY = F(u - F(F(F(2) * epxn) * rescale)), F(v - F(F(F(2) * epyn) * rescale))
        lastX, lastY = F(u + F(F(F(2) * epxn) * rescale)), F(v + F(F(F(2) * epyn) * rescale))
    if (firstX <= 0 or firstX >= width - 2 or firstY <= 0 or firstY >= height - 2
            or lastX <= 0 or lastX >= width - 2 or lastY <= 0 or lastY >= height - 2):
        return -1, None, None, None
    if not (rescale > F(0.7) and rescale < F(1.4)):
        return -1, None, None, None
    ex, ey = F(epxn * rescale), F(epyn * rescale)
    e2x, e2y = F(F(F(2) * epxn) * rescale), F(F(F(2) * epyn) * rescale)
    realVal_p1 = interp(kf_image, F(u + ex), F(v + ey))
    realVal_m1 = interp(kf_image, F(u - ex), F(v - ey))
    realVal = interp(kf_image, u, v)
    realVal_m2 = interp(kf_image, F(u - e2x), F(v - e2y))
    realVal_p2 = interp(kf_image, F(u + e2x), F(v + e2y))
    with np.errstate(all="ignore"):
        pClose = [F(pInf[i] + F(Kt[i] * max_idepth)) for i in range(3)]
        if pClose[2] < F(0.001):
            max_idepth = F(F(F(0.001) - pInf[2]) / Kt[2])
            pClose = [F(pInf[i] + F(Kt[i] * max_idepth)) for i in range(3)]
        z = pClose[2]
        pClose = [F(c / z) for c in pClose]
        pFar = [F(pInf[i] + F(Kt[i] * min_idepth)) for i in range(3)]
        if pFar[2] < F(0.001) or max_idepth < min_idepth:
            return -1, None, None, None
        z = pFar[2]
        pFar = [F(c / z) for c in pFar]
        if np.isnan(F(pFar[0] + pClose[0])):
            return -4, None, None, None
        incx, incy = F(pClose[0] - pFar[0]), F(pClose[1] - pFar[1])
        eplLength = F(np.sqrt(F(F(incx * incx) + F(incy * incy))))
        if eplLength == 0 or np.isinf(eplLength):               # `!eplLength > 0` is (!eplLength) > 0
            return -4, None, None, None
        if eplLength > MAX_EPL_LENGTH_CROP:
            pClose[0] = F(pFar[0] + F(F(incx * MAX_EPL_LENGTH_CROP) / eplLength))
            pClose[1] = F(pFar[1] + F(F(incy * MAX_EPL_LENGTH_CROP) / eplLength))
        incx = F(incx * F(GRADIENT_SAMPLE_DIST / eplLength))
        incy = F(incy * F(GRADIENT_SAMPLE_DIST / eplLength))
    pFar[0], pFar[1] = F(pFar[0] - incx), F(pFar[1] - incy)
    pClose[0], pClose[1] = F(pClose[0] + incx), F(pClose[1] + incy)
    if eplLength < MIN_EPL_LENGTH_CROP:
        pad = F(F(MIN_EPL_LENGTH_CROP - eplLength) / F(2.0))
        pFar[0], pFar[1] = F(pFar[0] - F(incx * pad)), F(pFar[1] - F(incy * pad))
        pClose[0], pClose[1] = F(pClose[0] + F(incx * pad)), F(pClose[1] + F(incy * pad))
    B = SAMPLE_POINT_TO_BORDER
    if pFar[0] <= B or pFar[0] >= width - B or pFar[1] <= B or pFar[1] >= height - B:
        return -1, None, None, None
    if pClose[0] <= B or pClose[0] >= width - B or pClose[1] <= B or pClose[1] >= height - B:
        with np.errstate(all="ignore"):
            if pClose[0] <= B:
                toAdd = F(F(F(B) - pClose[0]) / incx)
                pClose[0], pClose[1] = F(pClose[0] + F(toAdd * incx)), F(pClose[1] + F(toAdd * incy))
            elif pClose[0] >= width - B:
                toAdd = F(F(F(width - B) - pClose[0]) / incx)
                pClose[0], pClose[1] = F(pClose[0] + F(toAdd * incx)), F(pClose[1] + F(toAdd * incy))
            if pClose[1] <= B:
                toAdd = F(F(F(B) - pClose[1]) / incy)
                pClose[0], pClose[1] = F(pClose[0] + F(toAdd * incx)), F(pClose[1] + F(toAdd * incy))
            elif pClose[1] >= height - B:
                toAdd = F(F(F(height - B) - pClose[1]) / incy)
                pClose[0], pClose[1] = F(pClose[0] + F(toAdd * incx)), F(pClose[1] + F(toAdd * incy))
            fincx, fincy = F(pClose[0] - pFar[0]), F(pClose[1] - pFar[1])
            newLen = F(np.sqrt(F(F(fincx * fincx) + F(fincy * fincy))))
        if pClose[0] <= B or pClose[0] >= width - B or pClose[1] <= B or pClose[1] >= height - B or newLen < F(8.0):
            return -1, None, None, None
    cpx, cpy = pFar[0], pFar[1]
    rimg = ref.image
    v_m2 = interp(rimg, F(cpx - F(F(2.0) * incx)), F(cpy - F(F(2.0) * incy)))
    v_m1 = interp(rimg, F(cpx - incx), F(cpy - incy))
    v_0 = interp(rimg, cpx, cpy)
    v_p1 = interp(rimg, F(cpx + incx), F(cpy + incy))
    loop = 0
    bx = by = F(-1)
    best, second = INF, INF
    errPre = errPost = diffPre = diffPost = NAN
    bestWasLast = False
    eeLast = F(-1)
    eA = [NAN] * 5
    eB = [NAN] * 5
    cBest = cSecond = -1
    real = [realVal_p2, realVal_p1, realVal, realVal_m1, realVal_m2]
    with np.errstate(all="ignore"):
        while (((incx < 0) == (cpx > pClose[0])) and ((incy < 0) == (cpy > pClose[1]))) or loop == 0:
            v_p2 = interp(rimg, F(cpx + F(F(2) * incx)), F(cpy + F(F(2) * incy)))
            vals = [v_p2, v_p1, v_0, v_m1, v_m2]
            e = [F(vals[k] - real[k]) for k in range(5)]
            ee = F(0)
            for k in range(5):
                ee = F(ee + F(e[k] * e[k]))
            if loop % 2 == 0:
                eA = e
            else:
                eB = e

            def cross():
                s = F(eA[0] * eB[0])
                for k in range(1, 5):
                    s = F(s + F(eA[k] * eB[k]))
                return s
            if ee < best:
                second, cSecond = best, cBest
                best, cBest = ee, loop
                errPre = eeLast
                diffPre = cross()
                errPost = F(-1)
                diffPost = F(-1)
                bx, by = cpx, cpy
                bestWasLast = True
            else:
                if bestWasLast:
                    errPost = ee
                    diffPost = cross()
                    bestWasLast = False
                if ee < second:
                    second, cSecond = ee, loop
            eeLast = ee
            v_m2, v_m1, v_0, v_p1 = v_m1, v_0, v_p1, v_p2
            cpx, cpy = F(cpx + incx), F(cpy + incy)
            loop += 1
            if loop > 4096:
                raise RuntimeError("unbounded epipolar search")
        if best > F(4.0) * MAX_ERROR_STEREO:
            return -3, None, None, None
        if abs(cBest - cSecond) > 1.0 and F(MIN_DISTANCE_ERROR_STEREO * best) > second:
            return -2, None, None, None
        didSubpixel = False
        if G["useSubpixelStereo"]:
            gPre_pre = F(-(F(errPre - diffPre)))
            gPre_this = F(best - diffPre)
            gPost_this = F(-(F(best - diffPost)))
            gPost_post = F(errPost - diffPost)
            interpPost = interpPre = False
            if (gPost_this < 0) ^ (gPre_this < 0):
                pass
            elif (gPre_pre < 0) ^ (gPre_this < 0):
                if not ((gPost_post < 0) ^ (gPost_this < 0)):
                    interpPre = True
            elif (gPost_post < 0) ^ (gPost_this < 0):
                interpPost = True
            if interpPre:
                d = F(gPre_this / F(gPre_this - gPre_pre))
                bx, by = F(bx - F(d * incx)), F(by - F(d * incy))
                best = F(F(best - F(F(F(2) * d) * gPre_this)) - F(F(F(gPre_pre - gPre_this) * d) * d))
                didSubpixel = True
            elif interpPost:
                d = F(gPost_this / F(gPost_this - gPost_post))
                bx, by = F(bx + F(d * incx)), F(by + F(d * incy))
                best = F(F(best + F(F(F(2) * d) * gPost_this)) + F(F(F(gPost_post - gPost_this) * d) * d))
                didSubpixel = True
        sampleDist = F(GRADIENT_SAMPLE_DIST * rescale)
        gal = F(0)
        for a, b in ((realVal_p2, realVal_p1), (realVal_p1, realVal), (realVal, realVal_m1), (realVal_m1, realVal_m2)):
            t = F(a - b)
            gal = F(gal + F(t * t))
        gal = F(gal / F(sampleDist * sampleDist))
        if best > F(MAX_ERROR_STEREO + F(np.sqrt(gal) * F(20))):
            return -3, None, None, None
        oTt = ref.otherToThis_t
        if F(incx * incx) > F(incy * incy):
            oldX = F(F(fxi * bx) + cxi)
            nom = F(F(oldX * oTt[2]) - oTt[0])
            d0, d2 = dot3(KinvP, ref.row0), dot3(KinvP, ref.row2)
            idnew = F(F(d0 - F(oldX * d2)) / nom)
            alpha = F(F(F(incx * fxi) * F(F(d0 * oTt[2]) - F(d2 * oTt[0]))) / F(nom * nom))
        else:
            oldY = F(F(fyi * by) + cyi)
            nom = F(F(oldY * oTt[2]) - oTt[1])
            d1, d2 = dot3(KinvP, ref.row1), dot3(KinvP, ref.row2)
            idnew = F(F(d1 - F(oldY * d2)) / nom)
            alpha = F(F(F(incy * fyi) * F(F(d1 * oTt[2]) - F(d2 * oTt[1]))) / F(nom * nom))
        if idnew < 0 and not G["allowNegativeIdepths"]:
            return -2, None, None, None
        photo = F(F(F(4.0) * G["cameraPixelNoise2"]) / F(gal + DIVISION_EPS))
        tef = F(F(0.25) * F(F(1.0) + ref.initialTrackedResidual))
        g0, g1 = interp42(kf_grad, u, v)
        geo = F(F(F(g0 * epxn) + F(g1 * epyn)) + DIVISION_EPS)
        geo = F(F(F(tef * tef) * F(F(g0 * g0) + F(g1 * g1))) / F(geo * geo))
        var = F(F(alpha * alpha) * F(F(F(F(0.05) if didSubpixel else F(0.5)) * sampleDist * sampleDist + geo) + photo))
    return best, idnew, var, eplLength


def observe_pixel(cam, G, kf_image, kf_grad, kf_maxgrad, hyp, x, y, refs, state):
    """observeDepthRow's body for one pixel; `hyp` is a dict copy of the pixel's hypothesis, returned updated.
    refs: dict(oldest=, newest=, by_id=list, offset=int); state: dict(reactivated, numTracked, numMapped)"""
    h = dict(hyp)
    mg = kf_maxgrad[y, x]
    has = bool(h["isValid"])
    if has and mg < G["minUseGrad"]:
        h["isValid"] = 0
        return h
    if mg < G["minUseGrad"] or h["blacklisted"] < MIN_BLACKLIST:
        return h
    W1 = kf_image.shape[1] >> 1

    def tracked_badly(ref):
        return ref.good_mask is not None and not ref.good_mask[(y >> 1), (x >> 1)]
    if not has:                                                  # observeDepthCreate
        ref = refs["newest"] if state["reactivated"] else refs["oldest"]
        if tracked_badly(ref):
            return h
        ep = make_and_check_epl(cam, kf_image, x, y, ref)
        if ep is None:
            return h
        err, rid, rvar, rlen = do_line_stereo(cam, G, kf_image, kf_grad, F(x), F(y), ep[0], ep[1], F(0.0), F(1.0), F(F(1.0) / MIN_DEPTH), ref)
        if err == -3 or err == -2:
            h["blacklisted"] -= 1
        if err < 0 or rvar > MAX_VAR:
            return h
        h.update(isValid=1, blacklisted=0, nextStereoFrameMinID=F(0), validity_counter=5, idepth=unzero(rid), idepth_var=rvar,
                 idepth_smoothed=F(-1), idepth_var_smoothed=F(-1))
        return h
    # observeDepthUpdate
    if not state["reactivated"]:
        k = int(h["nextStereoFrameMinID"]) - refs["offset"]
        if k >= len(refs["by_id"]):
            return h
        ref = refs["oldest"] if k < 0 else refs["by_id"][k]
    else:
        ref = refs["newest"]
    if tracked_badly(ref):
        return h
    ep = make_and_check_epl(cam, kf_image, x, y, ref)
    if ep is None:
        return h
    with np.errstate(all="ignore"):
        sv = F(np.sqrt(h["idepth_var_smoothed"]))
        mn = F(h["idepth_smoothed"] - F(sv * STEREO_EPL_VAR_FAC))
        mx = F(h["idepth_smoothed"] + F(sv * STEREO_EPL_VAR_FAC))
    if mn < 0:
        mn = F(0)
    if mx > F(F(1) / MIN_DEPTH):
        mx = F(F(1) / MIN_DEPTH)
    err, rid, rvar, rlen = do_line_stereo(cam, G, kf_image, kf_grad, F(x), F(y), ep[0], ep[1], mn, h["idepth_smoothed"], mx, ref)
    if err == -1:
        return h
    if err == -2:
        h["validity_counter"] -= 5
        if h["validity_counter"] < 0:
            h["validity_counter"] = 0
        h["nextStereoFrameMinID"] = F(0)
        h["idepth_var"] = F(h["idepth_var"] * FAIL_VAR_INC_FAC)
        if h["idepth_var"] > MAX_VAR:
            h["isValid"] = 0
            h["blacklisted"] -= 1
        return h
    if err == -3 or err == -4:
        return h
    diff = F(rid - h["idepth_smoothed"])
    if F(F(F(1.0) * F(1.0)) * diff * diff) > F(rvar + h["idepth_var_smoothed"]):
        h["idepth_var"] = F(h["idepth_var"] * FAIL_VAR_INC_FAC)
        if h["idepth_var"] > MAX_VAR:
            h["isValid"] = 0
        return h
    id_var = F(h["idepth_var"] * SUCC_VAR_INC_FAC)
    w = F(rvar / F(rvar + id_var))
    new_idepth = F(F(F(F(1) - w) * rid) + F(w * h["idepth"]))
    h["idepth"] = unzero(new_idepth)
    id_var = F(id_var * w)
    if id_var < h["idepth_var"]:
        h["idepth_var"] = id_var
    h["validity_counter"] += 5
    cap = F(F(5.0) + F(F(mg * F(250.0)) / F(255.0)))
    if h["validity_counter"] > cap:
        h["validity_counter"] = int(cap)                           # float -> int member
    if rlen < MIN_EPL_LENGTH_CROP:
        inc = F(F(state["numTracked"]) / F(state["numMapped"] + 5))
        if inc < 3:
            inc = F(3)
        inc = F(inc + F(int(F(rlen * F(10000))) % 2))
        if D(rlen) < D(0.5) * D(MIN_EPL_LENGTH_CROP):
            inc = F(inc * F(3))
        h["nextStereoFrameMinID"] = F(F(ref.id) + inc)
    return h


def propagate_depth(cam, kf_image, new_image, new_maxgrad, hyp, new_to_old_qts, good_mask, min_use_grad=F(5.0)):
    """DepthMap::propagateDepth, DepthEstimation/DepthMap.cpp:475-653 (before the std::swap): -> the new map as a structured
    array.  new_to_old_qts = new_keyframe->pose->thisToParent_raw; good_mask = refPixelWasGoodNoCreate() or None."""
    H, W = hyp.shape
    out = hyp.copy()
    out["isValid"] = 0
    out["blacklisted"] = 0
    fx, fy, cx, cy = cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    fxi, fyi, cxi, cyi = cam["fxi"], cam["fyi"], cam["cxi"], cam["cyi"]
    # se3FromSim3 (util/SophusUtil.h:60-63: SE3(quaternion, translation), the constructor normalises) and .inverse() in double
    q = np.asarray(new_to_old_qts[:4], D)
    q = q / np.sqrt((q * q).sum())
    R = quat_R64(q).T
    t = -(R @ np.asarray(new_to_old_qts[4:7], D))
    Rf = R.astype(F)
    tf = np.array([F(v) for v in t], F)
    flat_old = kf_image.reshape(-1)
    ys, xs = np.nonzero(hyp["isValid"] > 0)
    with np.errstate(all="ignore"):
        for y, x in zip(ys.tolist(), xs.tolist()):
            src = hyp[y, x]
            ids = F(src["idepth_smoothed"])
            p = [F(F(F(x) * fxi) + cxi), F(F(F(y) * fyi) + cyi), F(1.0)]
            pn = [F(F(F(F(F(Rf[i, 0] * p[0]) + F(Rf[i, 1] * p[1])) + F(Rf[i, 2] * p[2])) / ids) + tf[i]) for i in range(3)]
            nid = F(F(1.0) / pn[2])
            u = F(F(F(pn[0] * nid) * fx) + cx)
            v = F(F(F(pn[1] * nid) * fy) + cy)
            if not (u > F(2.1) and v > F(2.1) and u < F(F(W) - F(3.1)) and v < F(F(H) - F(3.1))):
                continue
            nx, ny = int(F(u + F(0.5))), int(F(v + F(0.5)))
            dgrad = new_maxgrad[ny, nx]
            if good_mask is not None:
                if not good_mask[y >> 1, x >> 1] or dgrad < min_use_grad:
                    continue
            else:
                res = F(interp(new_image, u, v) - flat_old[x + y * W])
                if F(F(res * res) / F(F(40.0) * F(40.0) + F(F(F(0.5) * F(0.5)) * dgrad) * dgrad)) > F(1.0) or dgrad < min_use_grad:
                    continue
            tgt = out[ny, nx]
            r4 = F(nid / ids)
            r4 = F(r4 * r4)
            r4 = F(r4 * r4)
            nvar = F(r4 * F(src["idepth_var"]))
            valid = bool(tgt["isValid"])
            if valid:
                diff = F(F(tgt["idepth"]) - nid)
                if F(F(F(1.0) * F(1.0)) * diff * diff) > F(nvar + F(tgt["idepth_var"])):
                    if nid < F(tgt["idepth"]):
                        continue
                    valid = False
            if not valid:
                new = (nid, nvar, int(src["validity_counter"]))
            else:
                w = F(nvar / F(F(tgt["idepth_var"]) + nvar))
                merged = F(F(w * F(tgt["idepth"])) + F(F(F(1.0) - w) * nid))
                mv = int(src["validity_counter"]) + int(tgt["validity_counter"])
                if mv > F(5.0) + F(250.0):
                    mv = int(F(5.0) + F(250.0))
                new = (merged, F(F(1.0) / F(F(F(1.0) / F(tgt["idepth_var"])) + F(F(1.0) / nvar))), mv)
            out[ny, nx]["isValid"] = 1
            out[ny, nx]["blacklisted"] = 0
            out[ny, nx]["nextStereoFrameMinID"] = 0
            out[ny, nx]["validity_counter"] = new[2]
            out[ny, nx]["idepth"], out[ny, nx]["idepth_var"] = new[0], new[1]
            out[ny, nx]["idepth_smoothed"], out[ny, nx]["idepth_var_smoothed"] = -1, -1
    return out
