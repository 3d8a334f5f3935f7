"""image_to_3d / distort_keypoints of the C++ mirror (host code, no GPU): reference src/distort/distort_keypoints.cpp
:48-103, pinned the way test/test_distort.cpp pins it."""
import numpy as np

import oc_distort as D


def grid(cols, rows):
    return np.array([(i, j) for i in range(0, cols, cols // 20) for j in range(0, rows, rows // 20)], np.float64)


def test_no_distortion_is_bit_exact_and_round_trips(hostlib):
    # test/test_distort.cpp:34-43 -- f = 6000, 4000 x 3000, principal point at the centre; "inverts perfectly"
    cam = hostlib.camera8(6000, (2000, 1500))
    p = grid(4000, 3000)
    rays = hostlib.image_to_3d(p, cam)
    assert np.array_equal(rays, D.image_to_3d_undistorted(p, 6000, (2000, 1500)))
    assert np.allclose(np.linalg.norm(rays, axis=1), 1.0, atol=1e-15)
    assert np.abs(D.image_from_3d(rays, 6000, (2000, 1500)) - p).max() < 1e-12
    rng = np.random.default_rng(1)
    q = rng.uniform(-500, 6000, (5000, 2))
    assert np.array_equal(hostlib.image_to_3d(q, hostlib.camera8(5000, (2672, 2008))),
                          D.image_to_3d_undistorted(q, 5000, (2672, 2008)))


def test_distortion_round_trips_to_a_hundredth_of_a_pixel(hostlib):
    # test/test_distort.cpp:45-67 -- "only solve to 1/100 of a pixel error"
    p = grid(4000, 3000)
    for radial, tangential in (((0.02, -0.07, 0.1), (0, 0)), ((0.02, -0.07, 0.1), (0.08, -0.08)),
                               ((-0.05, 0, 0), (0, 0)), ((0, 0, 0), (0.01, 0.02))):
        cam = hostlib.camera8(6000, (2000, 1500), radial, tangential)
        rays = hostlib.image_to_3d(p, cam)
        assert np.allclose(np.linalg.norm(rays, axis=1), 1.0, atol=1e-14)
        back = D.image_from_3d(rays, 6000, (2000, 1500), radial, tangential)
        assert np.abs(back - p).max() < 1e-2, (radial, tangential, np.abs(back - p).max())
        # and the solve moved the point: the distorted and undistorted rays differ
        assert np.abs(rays - D.image_to_3d_undistorted(p, 6000, (2000, 1500))).max() > 1e-6
