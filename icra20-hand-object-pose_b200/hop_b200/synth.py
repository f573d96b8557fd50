"""Seeded synthetic inputs of the shapes named in BASELINE.json / SURVEY.md 8(d).

The reference's object models, PPF tables and hand meshes are not in its repository (external download), so every
test and benchmark runs on these stand-ins: analytic surfaces sampled uniformly by area with outward unit normals,
a partial camera-facing view with noise/outliers/confidence as the scene, and pose hypotheses scattered around the
ground truth.  Everything is float32 and a pure function of the seed.
"""
import numpy as np


def _rot_from_rotvec(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _ellipsoid(rng, n, a=0.045, b=0.030, c=0.020):
    pts, nrm = [], []
    gmax = max(b * c, a * c, a * b)
    need = n
    while need > 0:
        u = rng.normal(size=(2 * need + 64, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        g = np.sqrt((b * c * u[:, 0]) ** 2 + (a * c * u[:, 1]) ** 2 + (a * b * u[:, 2]) ** 2)
        keep = rng.random(len(u)) * gmax < g
        u = u[keep][:need]
        pts.append(u * [a, b, c])
        nn = u / [a, b, c]
        nrm.append(nn / np.linalg.norm(nn, axis=1, keepdims=True))
        need -= len(u)
    return np.concatenate(pts), np.concatenate(nrm)


def _cuboid(rng, n, sx=0.08, sy=0.05, sz=0.03):
    h = np.array([sx, sy, sz]) / 2
    areas = np.array([sy * sz, sy * sz, sx * sz, sx * sz, sx * sy, sx * sy])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    uv = rng.random((n, 3)) * 2 - 1
    pts = uv * h
    nrm = np.zeros((n, 3))
    ax, sg = face // 2, np.where(face % 2 == 0, 1.0, -1.0)
    pts[np.arange(n), ax] = sg * h[ax]
    nrm[np.arange(n), ax] = sg
    return pts, nrm


def _cylinder(rng, n, r=0.025, h=0.09, z0=0.0):
    a_side, a_cap = 2 * np.pi * r * h, np.pi * r * r
    kind = rng.choice(3, size=n, p=np.array([a_side, a_cap, a_cap]) / (a_side + 2 * a_cap))
    th = rng.random(n) * 2 * np.pi
    rad = r * np.sqrt(rng.random(n))
    z = (rng.random(n) - 0.5) * h
    pts = np.stack([np.where(kind == 0, r, rad) * np.cos(th), np.where(kind == 0, r, rad) * np.sin(th),
                    np.where(kind == 0, z, np.where(kind == 1, h / 2, -h / 2)) + z0], 1)
    nrm = np.stack([np.where(kind == 0, np.cos(th), 0.0), np.where(kind == 0, np.sin(th), 0.0),
                    np.where(kind == 0, 0.0, np.where(kind == 1, 1.0, -1.0))], 1)
    return pts, nrm


def _tless(rng, n):
    """cylinder r=0.03,h=0.04 with a coaxial flange r=0.045,h=0.01 at its base (T-LESS-like)."""
    p1, n1 = _cylinder(rng, 2 * n, 0.03, 0.04, 0.0)
    p2, n2 = _cylinder(rng, 2 * n, 0.045, 0.01, -0.025)
    in2 = (np.hypot(p1[:, 0], p1[:, 1]) < 0.045 - 1e-9) & (np.abs(p1[:, 2] + 0.025) < 0.005 - 1e-9)
    in1 = (np.hypot(p2[:, 0], p2[:, 1]) < 0.03 - 1e-9) & (np.abs(p2[:, 2]) < 0.02 - 1e-9)
    a1 = 2 * np.pi * 0.03 * 0.04 + 2 * np.pi * 0.03 ** 2
    a2 = 2 * np.pi * 0.045 * 0.01 + 2 * np.pi * 0.045 ** 2
    k1 = int(round(n * a1 / (a1 + a2)))
    p1, n1, p2, n2 = p1[~in2][:k1], n1[~in2][:k1], p2[~in1][: n - k1], n2[~in1][: n - k1]
    return np.concatenate([p1, p2]), np.concatenate([n1, n2])


MODELS = {"ellipse": _ellipsoid, "cuboid": _cuboid, "cylinder": _cylinder, "tless": _tless}


def make_model(name, n, seed=0):
    rng = np.random.default_rng(seed)
    pts, nrm = MODELS[name](rng, n)
    assert len(pts) == n
    return pts.astype(np.float32), nrm.astype(np.float32)


def make_gt_pose(rng):
    T = np.eye(4)
    T[:3, :3] = random_rotation(rng)
    T[:3, 3] = [rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), rng.uniform(0.30, 0.40)]
    return T


def make_scene(name, n_scene, seed=0, noise=0.0005, outlier_frac=0.10, dense_factor=6):
    """Partial view of the object under a random ground-truth pose.  Returns xyz, normals, confidence, gt_pose(4x4)."""
    rng = np.random.default_rng(seed)
    gt = make_gt_pose(rng)
    n_out = int(round(outlier_frac * n_scene))
    n_in = n_scene - n_out
    pts, nrm = MODELS[name](rng, max(dense_factor * n_in, 1024))
    P = pts @ gt[:3, :3].T + gt[:3, 3]
    N = nrm @ gt[:3, :3].T
    vis = np.einsum("ij,ij->i", N, P) < 0  # camera at the origin: keep surface facing it
    P, N = P[vis], N[vis]
    sel = rng.choice(len(P), size=n_in, replace=len(P) < n_in)
    P, N = P[sel], N[sel]
    P = P + N * rng.normal(0, noise, size=(n_in, 1))
    lo, hi = P.min(0) - 0.02, P.max(0) + 0.02
    Po = rng.uniform(lo, hi, size=(n_out, 3))
    No = rng.normal(size=(n_out, 3))
    No /= np.linalg.norm(No, axis=1, keepdims=True)
    xyz = np.concatenate([P, Po])
    nn = np.concatenate([N, No])
    perm = rng.permutation(n_scene)
    conf = rng.uniform(0.8, 1.0, size=n_scene)
    return xyz[perm].astype(np.float32), nn[perm].astype(np.float32), conf.astype(np.float32), gt.astype(np.float32)


def make_hypotheses(gt, H, seed=0, rot_sigma_deg=5.0, trans_sigma=0.005, random_frac=0.10):
    """H poses = GT o exp(xi) with xi_rot ~ N(0, 5 deg) per axis, xi_t ~ N(0, 5 mm); 10 % fully random."""
    rng = np.random.default_rng(seed)
    out = np.zeros((H, 4, 4), np.float64)
    for i in range(H):
        if rng.random() < random_frac:
            out[i] = make_gt_pose(rng)
        else:
            D = np.eye(4)
            D[:3, :3] = _rot_from_rotvec(rng.normal(0, np.deg2rad(rot_sigma_deg), size=3))
            D[:3, 3] = rng.normal(0, trans_sigma, size=3)
            out[i] = gt.astype(np.float64) @ D
    return out.astype(np.float32)


def pose_error(A, B):
    """translation error (m) and rotation error (deg) between batches of 4x4 poses."""
    A = np.asarray(A, np.float64).reshape(-1, 4, 4)
    B = np.asarray(B, np.float64).reshape(-1, 4, 4)
    dt = np.linalg.norm(A[:, :3, 3] - B[:, :3, 3], axis=1)
    # angle from the chord |Ra - Rb|_F = 2*sqrt(2)*sin(angle/2): well conditioned near 0 (arccos of the trace is not)
    chord = np.linalg.norm(A[:, :3, :3] - B[:, :3, :3], axis=(1, 2))
    return dt, np.rad2deg(2.0 * np.arcsin(np.clip(chord / (2.0 * np.sqrt(2.0)), 0.0, 1.0)))


SYMMETRY_AXIS = {"cylinder": 2, "tless": 2}  # continuous rotational symmetry about this model axis (through the origin)


def pose_error_sym(A, B, name):
    """pose_error that ignores the rotation about the object's continuous symmetry axis (it is unobservable: the
    reference folds it away the same way, object_symmetry in config_autodataset.yaml / PoseEstimator.cpp:134-190)."""
    dt, dr = pose_error(A, B)
    ax = SYMMETRY_AXIS.get(name)
    if ax is None:
        return dt, dr
    A = np.asarray(A, np.float64).reshape(-1, 4, 4)
    B = np.asarray(B, np.float64).reshape(-1, 4, 4)
    za, zb = A[:, :3, ax], B[:, :3, ax]
    za = za / np.linalg.norm(za, axis=1, keepdims=True)
    zb = zb / np.linalg.norm(zb, axis=1, keepdims=True)
    chord = np.linalg.norm(za - zb, axis=1)
    return dt, np.rad2deg(2.0 * np.arcsin(np.clip(chord / 2.0, 0.0, 1.0)))


def workload(name):
    """Named workloads (BASELINE.json configs, concretised in SURVEY.md 8d)."""
    table = {
        "C2": dict(model="ellipse", n_scene=2000, n_model=10000, H=1024, max_iter=10),
        "headline": dict(model="ellipse", n_scene=10000, n_model=10000, H=16384, max_iter=10),
        "C3": dict(model="cuboid", n_scene=2000, n_model=10000, H=8192, max_iter=50),
        "C5": dict(model="ellipse", n_scene=50000, n_model=50000, H=65536, max_iter=10),
        "tiny": dict(model="ellipse", n_scene=500, n_model=2000, H=64, max_iter=10),
    }
    return table[name]
