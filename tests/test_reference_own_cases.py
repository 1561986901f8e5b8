"""The known-answer cases of the REFERENCE'S OWN test-suite (tests/test_UtilsCV.py of Sahar-E/NeRF-and-DietNeRF, its only
test module) run against this framework's implementations of the same functions (pose geometry on the host:
``poses.py`` / ``UtilsFiles.py``).  Same inputs and expected values as the reference's cases (:15-186), table-driven."""
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def mods():
    pkg = importlib.import_module("nerf-and-dietnerf_b200")
    return pkg.poses, pkg.UtilsFiles


def test_normalize_vectors(mods):                                  # reference :15-33
    _, F = mods
    assert np.allclose(F.normalize_vectors(np.asarray([1, 1])), [0.7071, 0.7071], atol=1e-4)
    got = F.normalize_vectors(np.asarray([[1, 1], [1, 0], [0, 1]]))
    assert np.allclose(got, [[0.7071, 0.7071], [1, 0], [0, 1]], atol=1e-4)


def test_least_squares_intersection_of_lines(mods):                 # reference :35-54
    P, _ = mods
    dirs = np.asarray([[1, 1], [1, 1], [1, 1], [1, 0], [0, 1]], dtype=np.float64)
    points = np.asarray([[0, 0], [0, 0], [0, 0], [0, 1], [1, 0]], dtype=np.float64)
    assert np.allclose(P.estimate_intersection_between_lines(np.stack([dirs, points], axis=1)), [1, 1])


def _shifted(P, x, y, z, shift):
    m = P.get_sphere_matrix(1, x, y, z)
    m[:3, 3] += shift
    return m


@pytest.mark.parametrize("poses,expected,found", [
    ([(0, 0, 0, 0), (0, 90, 0, 0)], 0.0, True),                                                             # :56-62
    ([(0, 0, 0, 0), (90, 0, 0, 0), (0, 90, 0, 0), (0, 0, 90, 0), (0, 0, 90, 0), (0, 90, 0, 1), (0, 90, 0, -1)], 0.0, True),   # :64-89
    ([(0, 0, 0, 1), (90, 0, 0, 1), (0, 90, 0, 1), (0, 0, 90, 1), (0, 0, 90, 1), (0, 90, 0, 0), (0, 90, 0, 0)], 1.0, True),    # :91-118
    ([(0, 0, 0, 1), (90, 0, 0, -1), (0, 90, 0, 0)], None, False),                                            # :120-133
])
def test_point_of_interest_of_a_scene(mods, poses, expected, found):
    """RANSAC over the optical axes: cameras on a sphere (with two translated outliers) look at the sphere's centre;
    three cameras without a common point do not."""
    P, _ = mods
    c2ws = [_shifted(P, x, y, z, s) for x, y, z, s in poses]
    point, is_there = P.estimate_point_of_interest_in_scene(c2ws, rng=np.random.RandomState(0))
    assert is_there == found
    if found:
        assert np.allclose(point, np.full(3, expected), atol=1e-6)


def test_quaternion_between_two_vectors(mods):                      # reference :135-142
    P, _ = mods
    v1, v2 = np.asarray([1.0, 0, 0]), np.asarray([0, 1 / np.sqrt(2), 1 / np.sqrt(2)])      # a quarter turn apart
    q = P.get_rotation_quaternion_from_vec1_to_vec2(v1, v2)
    assert np.allclose(q, [1 / np.sqrt(2), 0, -0.5, 0.5])            # (w, x, y, z)
    assert np.allclose(P.rotate_vec_with_quaternion(v1, q), v2)


def test_camera_direction(mods):                                    # reference :144-147
    P, _ = mods
    assert np.allclose(P.get_camera_dir_from_c2w(P.get_sphere_matrix(1, 90, 0, 0)), [0, 1, 0])


@pytest.mark.parametrize("a,b", [((90, 0, 0), (0, 0, 0)), ((45, 0, 0), (0, 0, 0)), ((33, 133, 33), (5, 243, 12))])
def test_rotation_matrix_between_camera_directions(mods, a, b):     # reference :149-160, :169-186
    P, _ = mods
    v1 = P.get_camera_dir_from_c2w(P.get_sphere_matrix(1, *a))
    v2 = P.get_camera_dir_from_c2w(P.get_sphere_matrix(1, *b))
    rotation = P.get_rotation_matrix_from_v1_to_v2(v1, v2)
    assert np.allclose(rotation @ v1, v2)
    assert np.allclose(P.rotate_vec_with_quaternion(v1, P.get_rotation_quaternion_from_vec1_to_vec2(v1, v2)), v2)
    if a == (90, 0, 0):
        v3 = P.get_camera_dir_from_c2w(P.get_sphere_matrix(1, 0, 90, 0))        # perpendicular to both: the rotation axis
        assert np.allclose(rotation @ v3, v3)


def test_rotation_matrix_between_plain_vectors(mods):               # reference :162-167
    P, _ = mods
    v1, v2 = np.asarray([1.0, 0, 0]), np.asarray([0, 1 / np.sqrt(2), 1 / np.sqrt(2)])
    assert np.allclose(P.get_rotation_matrix_from_v1_to_v2(v1, v2) @ v1, v2)
