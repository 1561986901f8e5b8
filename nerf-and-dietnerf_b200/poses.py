"""Camera-pose geometry on the host (NumPy, 4x4 matrices): what DietNeRF's random source pose needs
(src/DietNeRF.py:238-259: a pose on a sphere looking at the origin, src/UtilsCV.py:41-121, and the slerp/lerp
interpolation between two camera-to-world matrices, :175-225), the RANSAC look-at point (:333-464) and the video
trajectories of ``ExecutionRun.render_video``'s callers (:146-158, :229-247, :407-437, :612-697).  The reference
leans on the ``quaternion`` package for the last group; the same algebra is written out here in (x, y, z, w) order.
"""
import numpy as np


def get_x_rot_mat(x_deg):
    a = np.deg2rad(x_deg)
    return np.asarray([[1, 0, 0, 0], [0, np.cos(a), -np.sin(a), 0], [0, np.sin(a), np.cos(a), 0], [0, 0, 0, 1]])


def get_y_rot_mat(y_deg):
    # the reference's sign convention (src/UtilsCV.py:86-98): -sin in the first row
    a = np.deg2rad(y_deg)
    return np.asarray([[np.cos(a), 0, -np.sin(a), 0], [0, 1, 0, 0], [np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]])


def get_z_rot_mat(z_deg):
    a = np.deg2rad(z_deg)
    return np.asarray([[np.cos(a), -np.sin(a), 0, 0], [np.sin(a), np.cos(a), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])


def get_sphere_matrix(radius, x_rot, y_rot, z_rot):
    """Camera at distance ``radius`` on the +z axis, rotated about x, then y, then z (degrees), looking at the origin
    (src/UtilsCV.py:101-121)."""
    t = np.eye(4)
    t[2, 3] = radius
    return get_z_rot_mat(z_rot) @ (get_y_rot_mat(y_rot) @ (get_x_rot_mat(x_rot) @ t))


def quaternion_from_rotation_matrix(r):
    """(x, y, z, w) unit quaternion of a 3x3 rotation matrix (branch on the largest diagonal term)."""
    r = np.asarray(r, dtype=np.float64)
    tr = r[0, 0] + r[1, 1] + r[2, 2]
    if tr > 0:
        s = 2.0 * np.sqrt(1.0 + tr)
        q = [(r[2, 1] - r[1, 2]) / s, (r[0, 2] - r[2, 0]) / s, (r[1, 0] - r[0, 1]) / s, 0.25 * s]
    elif r[0, 0] > r[1, 1] and r[0, 0] > r[2, 2]:
        s = 2.0 * np.sqrt(1.0 + r[0, 0] - r[1, 1] - r[2, 2])
        q = [0.25 * s, (r[0, 1] + r[1, 0]) / s, (r[0, 2] + r[2, 0]) / s, (r[2, 1] - r[1, 2]) / s]
    elif r[1, 1] > r[2, 2]:
        s = 2.0 * np.sqrt(1.0 + r[1, 1] - r[0, 0] - r[2, 2])
        q = [(r[0, 1] + r[1, 0]) / s, 0.25 * s, (r[1, 2] + r[2, 1]) / s, (r[0, 2] - r[2, 0]) / s]
    else:
        s = 2.0 * np.sqrt(1.0 + r[2, 2] - r[0, 0] - r[1, 1])
        q = [(r[0, 2] + r[2, 0]) / s, (r[1, 2] + r[2, 1]) / s, 0.25 * s, (r[1, 0] - r[0, 1]) / s]
    return np.asarray(q)


def rotation_matrix_from_quaternion(q):
    x, y, z, w = q
    return np.asarray([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def slerp(p0, p1, t):
    """src/UtilsCV.py:228-247: spherical interpolation of two unit quaternions along the shorter arc."""
    cos_a = float(np.dot(p0, p1))
    if cos_a < 0:
        p1, cos_a = -p1, -cos_a
    omega = np.arccos(min(cos_a, 1.0))
    sin_omega = np.sin(omega)
    if sin_omega < 1e-8:           # identical rotations: the reference divides 0/0 here; keep the rotation
        return p0
    return np.sin((1.0 - t) * omega) / sin_omega * p0 + np.sin(t * omega) / sin_omega * p1


def interpolation_type_slerp_for_c2w(c2w1, c2w2, alpha):
    """Slerp the rotations and lerp the translations of two camera-to-world matrices (src/UtilsCV.py:175-210);
    ``alpha`` a float or a sequence of floats (then a list is returned)."""
    def one(m1, m2, t):
        m1, m2 = np.asarray(m1, dtype=np.float64), np.asarray(m2, dtype=np.float64)
        q = slerp(quaternion_from_rotation_matrix(m1[:3, :3]), quaternion_from_rotation_matrix(m2[:3, :3]), t)
        out = np.eye(4)
        out[:3, :3] = rotation_matrix_from_quaternion(q / np.linalg.norm(q))
        out[:3, 3] = m1[:3, 3] * (1 - t) + m2[:3, 3] * t
        return out.astype(np.float32)
    alpha = np.asarray(alpha)
    if alpha.shape != ():
        return [one(c2w1, c2w2, float(a)) for a in alpha]
    return one(c2w1, c2w2, float(alpha))


# ---- scene point of interest (src/UtilsCV.py:333-404, :440-464), used by ExecutionRun._init_dietnerf --------------------
def _normalize(x):
    return x / np.linalg.norm(x, axis=-1)[..., None]


def get_camera_dir_from_c2w(c2w):
    """The camera looks down its -z axis (src/UtilsCV.py:602-609)."""
    return _normalize(-np.asarray(c2w)[:3, 2])


def estimate_intersection_between_lines(dirs_and_t):
    """Least-squares point closest to a set of lines given as (direction, point) pairs."""
    if dirs_and_t.shape[0] == 1:
        return None
    dirs, t = _normalize(dirs_and_t[:, 0]), dirs_and_t[:, 1]
    proj = np.eye(dirs.shape[-1]) - dirs[..., None] @ dirs[..., None, :]
    left = np.concatenate(proj, axis=0)
    right = np.concatenate(np.squeeze(proj @ t[..., None], -1), axis=0)
    return np.linalg.lstsq(left, right, rcond=None)[0]


def get_distance_of_point_from_line(point, dirs_and_t):
    """SQUARED distance of ``point`` from every line (what the reference compares with its tolerance)."""
    dirs, t = _normalize(dirs_and_t[:, 0]), dirs_and_t[:, 1]
    proj = np.eye(dirs.shape[-1]) - dirs[..., None] @ dirs[..., None, :]
    diff = (t - point)
    return np.squeeze(diff[..., None, :] @ proj @ diff[..., None], (-1, -2))


def ransac_get_estimation_for_intersection_point(dirs_and_t, num_iter=10000, inlier_tol=0.001, n_lines=2, rng=None):
    rng = np.random if rng is None else rng
    best_n, best_idx = -1, None
    for _ in range(num_iter):
        choice = rng.choice(dirs_and_t.shape[0], n_lines, replace=False)
        point = estimate_intersection_between_lines(dirs_and_t[choice])
        inlier = get_distance_of_point_from_line(point, dirs_and_t) < inlier_tol
        if inlier.sum() > best_n:
            best_n, best_idx = int(inlier.sum()), np.where(inlier)[0]
    if best_n > 1:
        point = estimate_intersection_between_lines(dirs_and_t[best_idx])
        return point, np.where(get_distance_of_point_from_line(point, dirs_and_t) < inlier_tol)[0]
    return None, None


def estimate_point_of_interest_in_scene(c2w_matrices, num_iter=10000, rng=None):
    """-> (estimated look-at point or None, is_spherical_dataset): spherical when more than 30 % of the optical axes
    pass within the tolerance of the estimate."""
    assert len(c2w_matrices) > 1
    dirs_and_t = np.asarray([[get_camera_dir_from_c2w(c), np.asarray(c)[:3, 3]] for c in c2w_matrices], dtype=np.float64)
    point, inliers = ransac_get_estimation_for_intersection_point(dirs_and_t, num_iter=num_iter, rng=rng)
    if point is not None and inliers is not None:
        return point, bool(inliers.shape[0] > 0.3 * dirs_and_t.shape[0])
    return None, False


def get_l_to_r_c2w_matrices(total_frames):
    """Left-to-right dolly looking forward (src/UtilsCV.py:407-426)."""
    mats = np.tile(np.eye(4, dtype=np.float32), (total_frames, 1, 1))
    mats[:, 0, 3] = np.linspace(0, 1, total_frames) * 2 - 1
    return mats


def get_sphere_matrices(total_n_matrices):
    """Orbit about y then about x on the unit sphere (src/UtilsCV.py:429-437): 2 * total_n_matrices poses."""
    mats = [get_sphere_matrix(1, 0, deg, 0) for deg in np.linspace(0, 360, total_n_matrices)] + \
           [get_sphere_matrix(1, deg, 0, 0) for deg in np.linspace(0, 360, total_n_matrices)]
    return np.asarray(mats, dtype=np.float32)


# ---- video trajectories between dataset poses (src/UtilsCV.py:146-158, :229-247) -----------------------------------------
def get_c2w_matrices_between_2_c2w(c2w1, c2w2, n_renders=16):
    """``n_renders`` poses from ``c2w1`` to ``c2w2`` at evenly spaced interpolation weights (both ends included)."""
    return interpolation_type_slerp_for_c2w(c2w1, c2w2, np.linspace(0, 1, n_renders))


def get_c2w_matrices_between_2_c2w_with_stretch(c2w1, c2w2, n_renders, stretch_knob=1):
    """As above with the weights warped by a/(a + 1 + knob) and re-normalised to [0, 1]: the camera starts fast and
    slows down towards ``c2w2`` (a larger knob stretches less)."""
    alpha = np.linspace(0, 1, n_renders)
    stretched = alpha * (1 / (alpha + 1 + stretch_knob))
    stretched = (stretched - stretched.min()) / (stretched.max() - stretched.min())
    return interpolation_type_slerp_for_c2w(c2w1, c2w2, stretched)


# ---- rotations between vectors / frames (src/UtilsCV.py:612-697) ----------------------------------------------------------
X_UNIT_VEC = np.asarray([1.0, 0.0, 0.0])
Y_UNIT_VEC = np.asarray([0.0, 1.0, 0.0])


def _quat_mul(a, b):
    """Hamilton product of two (w, x, y, z) quaternions."""
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.asarray([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                       aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def get_rotation_quaternion_with_axis_vec_and_theta(axis_vec, theta):
    """(w, x, y, z) rotation by ``theta`` radians about the unit vector ``axis_vec`` (the reference's order here)."""
    return np.concatenate(([np.cos(theta / 2)], np.asarray(axis_vec, dtype=np.float64) * np.sin(theta / 2)))


def get_rotation_quaternion_from_vec1_to_vec2(v1, v2):
    """(w, x, y, z) rotation taking the direction of ``v1`` onto that of ``v2``; anti-parallel inputs turn by pi about
    an axis orthogonal to them, parallel ones give the identity."""
    n1, n2 = _normalize(np.asarray(v1, dtype=np.float64)), _normalize(np.asarray(v2, dtype=np.float64))
    dot = float(n1.dot(n2))
    if dot < -0.99999:
        axis = np.cross(X_UNIT_VEC, n1)
        if np.linalg.norm(axis) < 0.00001:
            axis = np.cross(Y_UNIT_VEC, n1)
        return get_rotation_quaternion_with_axis_vec_and_theta(_normalize(axis), np.pi)
    if dot > 0.99999:
        return np.asarray([1.0, 0.0, 0.0, 0.0])
    return get_rotation_quaternion_with_axis_vec_and_theta(_normalize(np.cross(n1, n2)), np.arccos(dot))


def rotate_vec_with_quaternion(vec, q):
    """q * vec * q^-1 for a unit (w, x, y, z) quaternion."""
    q = np.asarray(q, dtype=np.float64)
    q_inv = np.asarray([q[0], -q[1], -q[2], -q[3]])
    return _quat_mul(_quat_mul(q, np.concatenate(([0.0], np.asarray(vec, dtype=np.float64)))), q_inv)[1:]


def get_rotation_matrix_from_v1_to_v2(v1, v2):
    w, x, y, z = get_rotation_quaternion_from_vec1_to_vec2(v1, v2)
    return rotation_matrix_from_quaternion((x, y, z, w))


def get_rotation_matrix_from_source_to_dest_mats(source_mat, dest_mat):
    """Homogeneous 4x4 R with R[:3,:3] @ source = dest, through q_dest * q_source^-1 like the reference (for proper
    rotations this is dest @ source^T; the quaternion round trip also projects slightly non-orthonormal inputs)."""
    qs, qd = quaternion_from_rotation_matrix(source_mat), quaternion_from_rotation_matrix(dest_mat)
    qs, qd = qs / np.linalg.norm(qs), qd / np.linalg.norm(qd)
    to_wxyz = lambda q: np.asarray([q[3], q[0], q[1], q[2]])
    qs_inv = to_wxyz(qs) * np.asarray([1.0, -1.0, -1.0, -1.0])
    w, x, y, z = _quat_mul(to_wxyz(qd), qs_inv)
    out = np.eye(4)
    out[:3, :3] = rotation_matrix_from_quaternion((x, y, z, w))
    return out
