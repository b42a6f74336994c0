"""Shared helpers for the parity tests: drive the oracle and the CUDA path through the same call sequence."""
import numpy as np

IDENT = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)


def hyp_equal_report(a: np.ndarray, b: np.ndarray):
    """Compare two hypothesis fields (structured arrays).  Fields other than the flags only count where valid."""
    rep = {}
    va, vb = a["isValid"] > 0, b["isValid"] > 0
    rep["valid_mismatch"] = int((va != vb).sum())
    rep["blacklist_mismatch"] = int((a["blacklisted"] != b["blacklisted"]).sum())
    both = va & vb
    for f in ("idepth", "idepth_var", "idepth_smoothed", "idepth_var_smoothed", "nextStereoFrameMinID"):
        x, y = a[f][both], b[f][both]
        rep[f + "_bitdiff"] = int((x.view(np.uint32) != y.view(np.uint32)).sum())
        den = np.maximum(np.abs(x), 1e-12)
        rep[f + "_maxrel"] = float(np.max(np.abs(x - y) / den)) if x.size else 0.0
    rep["validity_mismatch"] = int((a["validity_counter"][both] != b["validity_counter"][both]).sum())
    rep["n_valid"] = int(both.sum())
    return rep


def pose_err(qt_a, qt_b):
    """(relative translation error, rotation angle error in rad) between two (q,t) poses."""
    qa, qb = np.asarray(qt_a[:4]), np.asarray(qt_b[:4])
    ta, tb = np.asarray(qt_a[4:7]), np.asarray(qt_b[4:7])
    dt = np.linalg.norm(ta - tb) / max(np.linalg.norm(tb), 1e-12)
    d = abs(float(np.dot(qa, qb)))
    ang = 2 * np.arccos(min(1.0, d))
    return dt, ang


def rot_angle(qt):
    return 2 * np.arccos(min(1.0, abs(float(qt[3]))))
