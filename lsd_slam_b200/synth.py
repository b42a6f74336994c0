"""Deterministic synthetic grayscale sequences for the LSD-SLAM hot path.

A textured height-field ``Z(X, Y) = 2.0 + 0.5 sin(1.5 X) cos(1.2 Y)`` is ray-cast from a camera that
moves with a constant twist ``xi`` per frame (``T_k = exp(k xi)``, camera-k-from-world, Sophus tangent
order [upsilon | omega], thirdparty/Sophus/sophus/se3.hpp:395-397).  Every frame, the keyframe
included, is produced the same way, so frame 0's z-depth is an exact ground-truth depth map for the
reference's ``gtDepthInit`` path (SlamSystem.cpp:831-854).

Only numpy + scipy.ndimage are used; everything is seeded, nothing is read from disk.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

DEFAULT_XI = np.array([0.004, 0.001, 0.0005, 0.0004, 0.0008, 0.0002], dtype=np.float64)


def pinhole_K(w: int, h: int) -> np.ndarray:
    """fx = fy = 525 * W/640, principal point at the image centre (calib/pinhole_example_calib.cfg style)."""
    f = 525.0 * w / 640.0
    return np.array([[f, 0, (w - 1) / 2.0], [0, f, (h - 1) / 2.0], [0, 0, 1]], dtype=np.float32)


def se3_exp(xi: np.ndarray):
    """exp of a twist [upsilon | omega] -> (R, t), double precision (closed form of se3.hpp:406-428)."""
    ups, om = np.asarray(xi[:3], np.float64), np.asarray(xi[3:], np.float64)
    th = np.linalg.norm(om)
    Om = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if th < 1e-10:
        R = np.eye(3) + Om
        V = np.eye(3) + 0.5 * Om
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th**2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * Om + (th - np.sin(th)) / th**3 * Om @ Om
    return R, V @ ups


def rot_to_quat(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> unit quaternion (x, y, z, w)."""
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


def quat_to_rot(q: np.ndarray) -> np.ndarray:
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_texture(w: int, h: int, seed: int = 1234) -> np.ndarray:
    """Blob-edge texture at 2W x 2H (periodic), float32, range about [64, 194]."""
    rng = np.random.default_rng(seed)
    n1 = rng.random((2 * h, 2 * w), dtype=np.float32)
    n2 = rng.random((2 * h, 2 * w), dtype=np.float32)
    bg = ndimage.gaussian_filter(n1, 60.0, mode="wrap")
    bg = (bg - bg.min()) / (bg.max() - bg.min())
    bl = ndimage.gaussian_filter(n2, 12.0, mode="wrap")
    bl = (bl > np.median(bl)).astype(np.float32)
    bl = ndimage.gaussian_filter(bl, 2.4, mode="wrap")
    return (64.0 + 40.0 * bg + 90.0 * bl).astype(np.float32)


def _height(X, Y):
    return 2.0 + 0.5 * np.sin(1.5 * X) * np.cos(1.2 * Y)


def semi_dense_fraction(image_u8: np.ndarray, min_use_grad: float = 5.0) -> float:
    """Fraction of pixels with maxGradients >= minUseGrad, with the reference's definitions: central-difference gradient
    (Frame.cpp:658-677) and the 3x1 / 1x3 maximum filters over LINEAR index ranges (Frame.cpp:708-759).  SURVEY 8d asks the
    generator to report this and to refuse streams outside 30-50 % (a texture that is dense everywhere is unrepresentative)."""
    img = image_u8.astype(np.float32)
    h, w = img.shape
    n = w * h
    f = img.reshape(n)
    a = np.zeros(n, np.float32)
    i = np.arange(w, n - w)
    gx = np.float32(0.5) * (f[i + 1] - f[i - 1])
    gy = np.float32(0.5) * (f[i + w] - f[i - w])
    a[i] = np.sqrt(gx * gx + gy * gy)
    j = np.arange(w + 1, n - w - 1)
    t = np.zeros(n, np.float32)
    t[j] = np.maximum(np.maximum(a[j - w], a[j]), a[j + w])
    m = a.copy()
    m[j] = np.maximum(np.maximum(t[j - 1], t[j]), t[j + 1])
    return float((m >= np.float32(min_use_grad)).mean())


class Sequence:
    """A synthetic stream: ``frame(k)`` -> uint8 image, ``pose(k)`` -> camera-k-from-world (R, t)."""

    def __init__(self, w: int = 640, h: int = 480, seed: int = 1234, xi: np.ndarray = DEFAULT_XI, noise: float = 1.0):
        if w % 16 or h % 16:
            raise ValueError("width and height must be multiples of 16 (SlamSystem.cpp:55)")
        self.w, self.h, self.seed, self.noise = w, h, seed, noise
        self.xi = np.asarray(xi, np.float64)
        self.K = pinhole_K(w, h)
        self.tex = make_texture(w, h, seed)
        f, cx, cy = float(self.K[0, 0]), float(self.K[0, 2]), float(self.K[1, 2])
        u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        self._rays = np.stack([(u - cx) / f, (v - cy) / f, np.ones_like(u)], axis=-1)

    def check_density(self, lo: float = 0.30, hi: float = 0.50) -> float:
        """semi-dense density of frame 0; raises outside [lo, hi]"""
        frac = semi_dense_fraction(self.frame(0))
        if not lo <= frac <= hi:
            raise ValueError(f"synthetic stream has maxGrad >= 5 on {100 * frac:.1f} % of the pixels, outside {100 * lo:.0f}-{100 * hi:.0f} %")
        return frac

    def pose(self, k: int):
        return se3_exp(k * self.xi)

    def frame_to_ref_qt(self, k: int, ref: int = 0) -> np.ndarray:
        """Ground-truth frameToReference = T_ref * T_k^-1 as (qx,qy,qz,qw,tx,ty,tz)."""
        Rk, tk = self.pose(k)
        Rr, tr = self.pose(ref)
        R = Rr @ Rk.T
        t = tr - R @ tk
        return np.concatenate([rot_to_quat(R), t])

    def _cast(self, k: int):
        R, t = self.pose(k)
        c = -R.T @ t                           # camera centre in world
        r = self._rays @ R                     # world ray directions (R^T d), rows
        s = np.full(r.shape[:2], 2.0)
        for _ in range(10):
            X = c[0] + s * r[..., 0]
            Y = c[1] + s * r[..., 1]
            fz = c[2] + s * r[..., 2] - _height(X, Y)
            Zx = 0.75 * np.cos(1.5 * X) * np.cos(1.2 * Y)
            Zy = -0.6 * np.sin(1.5 * X) * np.sin(1.2 * Y)
            s = s - fz / (r[..., 2] - (Zx * r[..., 0] + Zy * r[..., 1]))
        return c[0] + s * r[..., 0], c[1] + s * r[..., 1], s

    def render(self, k: int):
        """-> (uint8 image HxW, float32 z-depth HxW in camera k)."""
        X, Y, s = self._cast(k)
        f, cx, cy = float(self.K[0, 0]), float(self.K[0, 2]), float(self.K[1, 2])
        tx = (X * f / 2.0 + cx) * 2.0
        ty = (Y * f / 2.0 + cy) * 2.0
        th, tw = self.tex.shape
        x0 = np.floor(tx)
        y0 = np.floor(ty)
        ax = (tx - x0).astype(np.float32)
        ay = (ty - y0).astype(np.float32)
        xi0 = np.mod(x0.astype(np.int64), tw)
        yi0 = np.mod(y0.astype(np.int64), th)
        xi1 = np.mod(xi0 + 1, tw)
        yi1 = np.mod(yi0 + 1, th)
        T = self.tex
        val = ((1 - ax) * (1 - ay) * T[yi0, xi0] + ax * (1 - ay) * T[yi0, xi1]
               + (1 - ax) * ay * T[yi1, xi0] + ax * ay * T[yi1, xi1])
        if self.noise > 0:
            rng = np.random.default_rng(self.seed + k)
            val = val + self.noise * rng.standard_normal(val.shape, dtype=np.float32)
        img = np.clip(np.rint(val), 0, 255).astype(np.uint8)
        return img, s.astype(np.float32)

    def frame(self, k: int) -> np.ndarray:
        return self.render(k)[0]
