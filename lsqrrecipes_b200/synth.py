"""Seeded synthetic inputs with the shapes of the reference's example generators.

Each generator returns (data [n, D] float64 in the host AoS layout of the reference's datum
type, true_params).  The layouts are the ones that cross the drop-in boundary:
Point<double,n> = n doubles (common/Point.h:127), std::pair<Point3D,Point3D> = 6 doubles,
Ray3D = p then n (common/Ray3D.h:23-24), Frame = rotation[3][3] row-major then translation[3]
(common/Frame.h:30-31; the trailing int outputFormat is dropped when packing).

Shapes follow examples/planeEstimation.cxx:153-200, examples/sphereEstimation.cxx:136-184,
examples/AbsoluteOrientation.cxx:39-88.  numpy only.
"""
import numpy as np

SEED = 20261017


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def plane(n, outlier_frac=0.4, seed=SEED, dim=3, sigma=0.4, coord_max=1000.0, outlier_dist=20.0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    normal = _unit(rng.uniform(0, 1, dim))
    a = rng.uniform(-coord_max, coord_max, dim)
    n_out = int(round(n * outlier_frac))
    n_in = n - n_out
    p = rng.uniform(-coord_max, coord_max, (n_in, dim))
    t = p - a
    p = a + rng.normal(0, sigma, (n_in, dim)) + (t - (t @ normal)[:, None] * normal)
    outs = []
    need = n_out
    while need > 0:
        q = rng.uniform(-coord_max, coord_max, (max(need * 2, 16), dim))
        q = q[np.abs((q - a) @ normal) >= outlier_dist][:need]
        outs.append(q)
        need -= len(q)
    data = np.concatenate([p] + outs) if outs else p
    data = data[rng.permutation(n)]
    return np.ascontiguousarray(data, dtype=dtype), np.concatenate([normal, a])


def line(n, dim=2, outlier_frac=0.3, seed=SEED, sigma=0.4, coord_max=1000.0, outlier_dist=20.0):
    rng = np.random.default_rng(seed)
    direction = _unit(rng.uniform(0, 1, dim))
    a = rng.uniform(-coord_max, coord_max, dim)
    n_out = int(round(n * outlier_frac))
    n_in = n - n_out
    s = rng.uniform(-coord_max, coord_max, n_in)
    p = a + s[:, None] * direction + rng.normal(0, sigma, (n_in, dim))
    outs = []
    need = n_out
    while need > 0:
        q = rng.uniform(-coord_max, coord_max, (max(need * 2, 16), dim))
        v = q - a
        dist = np.linalg.norm(v - (v @ direction)[:, None] * direction, axis=1)
        q = q[dist >= outlier_dist][:need]
        outs.append(q)
        need -= len(q)
    data = np.concatenate([p] + outs) if outs else p
    data = data[rng.permutation(n)]
    return np.ascontiguousarray(data), np.concatenate([direction, a])


def sphere(n, dim=3, outlier_frac=0.4, seed=SEED, sigma=0.4, coord_max=1000.0, outlier_dist=20.0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-coord_max, coord_max, dim)
    r = rng.uniform(100.0, coord_max)
    n_out = int(round(n * outlier_frac))
    n_in = n - n_out
    u = _unit(rng.uniform(-1, 1, (n_in, dim)))
    p = c + r * u + rng.normal(0, sigma, (n_in, dim))
    outs = []
    need = n_out
    while need > 0:
        q = rng.uniform(-coord_max, coord_max, (max(need * 2, 16), dim))
        q = q[np.abs(np.linalg.norm(q - c, axis=1) - r) >= outlier_dist][:need]
        outs.append(q)
        need -= len(q)
    data = np.concatenate([p] + outs) if outs else p
    data = data[rng.permutation(n)]
    return np.ascontiguousarray(data), np.concatenate([c, [r]])


def quat_to_matrix(s, qx, qy, qz):
    return np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - s * qz), 2 * (qx * qz + s * qy)],
                     [2 * (qx * qy + s * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - s * qx)],
                     [2 * (qx * qz - s * qy), 2 * (qy * qz + s * qx), 1 - 2 * (qx * qx + qy * qy)]])


def _random_quat(rng):
    qx = rng.uniform(0, 1)
    qy = rng.uniform(0, np.sqrt(1 - qx * qx))
    qz = rng.uniform(0, np.sqrt(max(1 - qx * qx - qy * qy, 0.0)))
    s = np.sqrt(max(1 - qx * qx - qy * qy - qz * qz, 0.0))
    return np.array([s, qx, qy, qz])


def absolute_orientation(n, outlier_frac=0.3, seed=SEED, sigma=1.0):
    rng = np.random.default_rng(seed)
    q = _random_quat(rng)
    R = quat_to_matrix(*q)
    t = rng.uniform(-1000, 1000, 3)
    p1 = rng.uniform(-100, 100, (n, 3))
    p2 = p1 @ R.T + t + rng.normal(0, sigma, (n, 3))
    n_out = int(round(n * outlier_frac))
    idx = rng.permutation(n)[:n_out]
    p2[idx] = rng.uniform(-1100, 1100, (n_out, 3))
    return np.ascontiguousarray(np.concatenate([p1, p2], axis=1)), np.concatenate([q, t])


def rays(n, outlier_frac=0.3, seed=SEED, sigma=0.3):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-100, 100, 3)
    p = rng.uniform(-1000, 1000, (n, 3))
    target = x + rng.normal(0, sigma, (n, 3))
    n_out = int(round(n * outlier_frac))
    idx = rng.permutation(n)[:n_out]
    target[idx] = rng.uniform(-1000, 1000, (n_out, 3))
    d = _unit(target - p)
    return np.ascontiguousarray(np.concatenate([p, d], axis=1)), x


def pivot_frames(n, outlier_frac=0.2, seed=SEED, sigma=0.2):
    """Tool poses pivoting about a fixed world point: R_i tDRF + t_i = tW."""
    rng = np.random.default_rng(seed)
    t_drf = rng.uniform(-200, 200, 3)
    t_w = rng.uniform(-1000, 1000, 3)
    out = np.zeros((n, 12))
    n_out = int(round(n * outlier_frac))
    bad = set(rng.permutation(n)[:n_out].tolist())
    for i in range(n):
        q = _unit(rng.normal(0, 1, 4))
        R = quat_to_matrix(*q)
        t = t_w - R @ t_drf + rng.normal(0, sigma, 3)
        if i in bad:
            t = t + rng.uniform(-50, 50, 3)
        out[i, :9] = R.ravel()
        out[i, 9:] = t
    return out, np.concatenate([t_drf, t_w])


def dense_rows(n, nc, outlier_frac=0.3, seed=SEED, sigma=0.03):
    """Augmented rows [a | b] of a consistent system a.x = b: coefficients and solution ~ U(-100, 100) as in
    examples/linearEquationSystemSolver.cxx:58-66, small noise on b, gross outliers in b."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-100, 100, nc)
    A = rng.uniform(-100, 100, (n, nc))
    b = A @ x + rng.normal(0, sigma, n)
    n_out = int(round(n * outlier_frac))
    idx = rng.permutation(n)[:n_out]
    b[idx] += rng.uniform(5, 5000, n_out) * rng.choice([-1.0, 1.0], n_out)
    return np.ascontiguousarray(np.concatenate([A, b[:, None]], axis=1)), x


def euler_zyx(oz, oy, ox):
    """R = Rz(oz) Ry(oy) Rx(ox), the parametrisation of SinglePointTargetUSCalibrationParametersEstimator.cxx:430-441."""
    cz, sz, cy, sy, cx, sx = np.cos(oz), np.sin(oz), np.cos(oy), np.sin(oy), np.cos(ox), np.sin(ox)
    return np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx],
                     [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx],
                     [-sy, cy * sx, cy * cx]])


def crosswire(n, outlier_frac=0.3, seed=SEED, sigma=1.0):
    """Cross-wire phantom US calibration data in the shape of testing/SinglePointTargetUSCalibration
    ParametersEstimatorTest.cxx:556-650: a fixed point t1 seen in n tracked images.  Datum = [R2 (9), t2 (3), u, v]
    with  R2 (T3 (u, v, 0, 1)) + t2 = t1;  pixel noise sigma, gross outliers in (u, v).
    Returns the data and the 20 reference parameters
    [t1, t3, omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1), m_y R3(:,2), R3(:,3)]."""
    rng = np.random.default_rng(seed)
    m_x, m_y = 0.143, 0.139
    ox, oy, oz = rng.uniform(0.2, 1.3, 3)           # away from the gimbal lock of the Euler extraction
    R3 = euler_zyx(oz, oy, ox)
    t3 = rng.uniform(-100, 100, 3)
    t1 = rng.uniform(-100, 100, 3)
    out = np.zeros((n, 14))
    n_out = int(round(n * outlier_frac))
    bad = np.zeros(n, dtype=bool)
    bad[rng.permutation(n)[:n_out]] = True
    for i in range(n):
        u, v = rng.uniform(0, 640), rng.uniform(0, 480)
        p = R3 @ np.array([m_x * u, m_y * v, 0.0]) + t3           # point in the US reference frame
        R2 = quat_to_matrix(*_unit(rng.normal(0, 1, 4)))
        t2 = t1 - R2 @ p
        un, vn = u + rng.normal(0, sigma), v + rng.normal(0, sigma)
        if bad[i]:
            un, vn = rng.uniform(0, 640), rng.uniform(0, 480)
        out[i, :9] = R2.ravel()
        out[i, 9:12] = t2
        out[i, 12:] = (un, vn)
    prm = np.concatenate([t1, t3, [oz, oy, ox, m_x, m_y], m_x * R3[:, 0], m_y * R3[:, 1], R3[:, 2]])
    return out, prm


def calibrated_pointer(n, outlier_frac=0.3, seed=SEED, sigma=1.0):
    """Calibrated-pointer US calibration data (testing/SinglePointTargetUSCalibrationParametersEstimatorTest.cxx,
    generateCalibratedPointerData): the pointer tip p_i (tracker coordinates) is seen at pixel (u, v) of image i.
    Datum = [R2 (9), t2 (3), u, v, p (3)] with R2 (T3 (u, v, 0, 1)) + t2 = p.  Returns the data and the 17 reference
    parameters [t3, omega_z, omega_y, omega_x, m_x, m_y, m_x R3(:,1), m_y R3(:,2), R3(:,3)]."""
    rng = np.random.default_rng(seed)
    m_x, m_y = 0.143, 0.139
    ox, oy, oz = rng.uniform(0.2, 1.3, 3)
    R3 = euler_zyx(oz, oy, ox)
    t3 = rng.uniform(-100, 100, 3)
    out = np.zeros((n, 17))
    n_out = int(round(n * outlier_frac))
    bad = np.zeros(n, dtype=bool)
    bad[rng.permutation(n)[:n_out]] = True
    for i in range(n):
        u, v = rng.uniform(0, 640), rng.uniform(0, 480)
        pus = R3 @ np.array([m_x * u, m_y * v, 0.0]) + t3
        R2 = quat_to_matrix(*_unit(rng.normal(0, 1, 4)))
        t2 = rng.uniform(-300, 300, 3)
        p = R2 @ pus + t2
        un, vn = u + rng.normal(0, sigma), v + rng.normal(0, sigma)
        if bad[i]:
            un, vn = rng.uniform(0, 640), rng.uniform(0, 480)
        out[i, :9] = R2.ravel()
        out[i, 9:12] = t2
        out[i, 12:14] = (un, vn)
        out[i, 14:] = p
    prm = np.concatenate([t3, [oz, oy, ox, m_x, m_y], m_x * R3[:, 0], m_y * R3[:, 1], R3[:, 2]])
    return out, prm


def frames_from_quat_file(path):
    """Rows 'x y z qx qy qz qs' (testing/Data/pivotCalibrationData.txt) -> packed frames,
    following testing/PivotCalibrationParametersEstimatorTest.cxx:29-33."""
    raw = np.loadtxt(path)
    out = np.zeros((len(raw), 12))
    for i, (x, y, z, qx, qy, qz, qs) in enumerate(raw):
        out[i, :9] = quat_to_matrix(qs, qx, qy, qz).ravel()
        out[i, 9:] = (x, y, z)
    return out


GENERATORS = {
    "plane3": lambda n, seed=SEED: plane(n, seed=seed),
    "line2d": lambda n, seed=SEED: line(n, 2, seed=seed),
    "line2": lambda n, seed=SEED: line(n, 2, seed=seed),
    "line3": lambda n, seed=SEED: line(n, 3, seed=seed),
    "circle2": lambda n, seed=SEED: sphere(n, 2, seed=seed),
    "sphere3": lambda n, seed=SEED: sphere(n, 3, seed=seed),
    "absor": lambda n, seed=SEED: absolute_orientation(n, seed=seed),
    "ray": lambda n, seed=SEED: rays(n, seed=seed),
    "pivot": lambda n, seed=SEED: pivot_frames(n, seed=seed),
    "dense5": lambda n, seed=SEED: dense_rows(n, 5, seed=seed),
    "dense6": lambda n, seed=SEED: dense_rows(n, 6, seed=seed),
    "sphere4": lambda n, seed=SEED: sphere(n, 4, seed=seed),
    "plane4": lambda n, seed=SEED: plane(n, seed=seed, dim=4),
    "usxw": lambda n, seed=SEED: crosswire(n, seed=seed),
    "uscp": lambda n, seed=SEED: calibrated_pointer(n, seed=seed),
}
# the wider template space: fewer outliers as the minimal subset grows (0.8^9 = 13 % all-inlier subsets for the 8-D sphere)
for _d in (2, 5, 6, 7, 8):
    GENERATORS[f"plane{_d}"] = lambda n, seed=SEED, _d=_d: plane(n, seed=seed, dim=_d, outlier_frac=0.4 if _d == 2 else 0.2)
for _d in (5, 6, 7, 8):
    GENERATORS[f"sphere{_d}"] = lambda n, seed=SEED, _d=_d: sphere(n, _d, seed=seed, outlier_frac=0.2)
for _d in (4, 5, 6, 7, 8):
    GENERATORS[f"line{_d}"] = lambda n, seed=SEED, _d=_d: line(n, _d, seed=seed, sigma=0.1)   # the distance to a line grows like sigma sqrt(d - 1)
for _n in (2, 3, 4, 7, 8):
    GENERATORS[f"dense{_n}"] = lambda n, seed=SEED, _n=_n: dense_rows(n, _n, seed=seed, outlier_frac=0.3 if _n <= 4 else 0.2)
DELTAS = {"plane3": 0.5, "line2d": 0.5, "line2": 0.5, "line3": 0.5, "circle2": 0.5, "sphere3": 0.5, "absor": 2.0, "ray": 1.0, "pivot": 1.0,
          "dense5": 0.2, "dense6": 0.2, "usxw": 1.0, "uscp": 1.0, "sphere4": 0.5, "plane4": 0.5}
for _name in GENERATORS:
    if _name not in DELTAS:
        DELTAS[_name] = 0.2 if _name.startswith("dense") else 0.5


def random_subsets(n, k, H, seed=SEED):
    """H ordered k-subsets of distinct indices (draw order, like RANSAC.hxx:56-68)."""
    rng = np.random.default_rng(seed + 7)
    s = rng.integers(0, n, size=(H, k), dtype=np.int64)
    for _ in range(64):
        srt = np.sort(s, axis=1)
        bad = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
        if not bad.any():
            break
        s[bad] = rng.integers(0, n, size=(int(bad.sum()), k))
    return np.ascontiguousarray(s, dtype=np.int32)


def degenerate_pool(name, n=48, seed=SEED):
    """A data pool whose random minimal subsets are often DEGENERATE the way the reference's estimate() bodies test for:
    repeated records, records that are exact copies up to a scale or a shift along one axis (collinear / coplanar / rank-
    deficient subsets), all-zero records -- next to ordinary data, so that valid and invalid subsets are mixed."""
    rng = np.random.default_rng(seed)
    data, _ = GENERATORS[name](n, seed=seed)
    d = data.shape[1]
    q = n // 4
    data[q:2 * q] = data[0]                                            # copies of one record
    if d > 8:                                                           # frames (rotation + translation + ...): only repetition keeps a record meaningful
        return np.ascontiguousarray(data[rng.permutation(n)])
    base, step = data[1].copy(), np.zeros(d)
    step[0] = 1.0
    for i in range(2 * q, 3 * q):                                      # records on one axis-parallel line through record 1
        data[i] = base + step * float(i - 2 * q)
    data[3 * q] = 0.0                                                   # an all-zero record
    data[3 * q + 1] = 2.0 * data[2]                                     # a scaled copy
    return np.ascontiguousarray(data[rng.permutation(n)])
