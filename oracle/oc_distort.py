"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the pixel <-> ray functions of the reference's src/distort
(src/distort/distort_keypoints.cpp:48-103, include/opencalibration/distort/distort_keypoints.hpp:27-60).
Only tests/ import this.

image_to_3d without distortion is plain IEEE arithmetic and is restated operation by operation (bit-exact with any
correct implementation of the same expression order). With distortion the reference inverts distortProjectedRay with
ceres::TinySolver (Ceres is external and un-vendored: "parity unpinned" at the bit level); the oracle for that branch
is the property the reference's own tests pin (test/test_distort.py:45-67): image_from_3d(image_to_3d(p)) == p to
1e-2 px, with the forward model restated here.
"""
import numpy as np


def image_to_3d_undistorted(xy, f, pp):
    """(kp - pp) / f -> homogeneous -> normalized (distort_keypoints.cpp:67,96-97), one rounding per operation."""
    xy = np.asarray(xy, np.float64).reshape(-1, 2)
    x = (xy[:, 0] - pp[0]) / f
    y = (xy[:, 1] - pp[1]) / f
    n = np.sqrt((x * x + y * y) + 1.0)
    return np.stack([x / n, y / n, 1.0 / n], axis=1)


def distort_projected_ray(p, radial, tangential):
    """distortProjectedRay (distort_keypoints.hpp:27-43)."""
    p = np.asarray(p, np.float64).reshape(-1, 2)
    x, y = p[:, 0], p[:, 1]
    r2 = x * x + y * y
    rad = 1.0 + (radial[0] * r2 + radial[1] * r2 * r2 + radial[2] * r2 * r2 * r2)
    xd = rad * x + 2 * x * y * tangential[0] + tangential[1] * (r2 + 2 * x * x)
    yd = rad * y + 2 * x * y * tangential[1] + tangential[0] * (r2 + 2 * y * y)
    return np.stack([xd, yd], axis=1)


def image_from_3d(rays, f, pp, radial=(0, 0, 0), tangential=(0, 0)):
    """Forward model (distort_keypoints.hpp:45-66): project (z clamped at 1e-3), distort, to pixels."""
    rays = np.asarray(rays, np.float64).reshape(-1, 3)
    z = np.maximum(rays[:, 2], 1e-3)
    proj = rays[:, :2] / z[:, None]
    return distort_projected_ray(proj, radial, tangential) * f + np.asarray(pp, np.float64)
