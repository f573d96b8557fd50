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
        # 128 depth frames x 4096 hypotheses over 8 GPUs = 16 frames per rank per step (weak scaling: every rank its own frames)
        "C4": dict(model="ellipse", n_scene=2000, n_model=10000, H=4096, max_iter=10, frames_per_step=16),
        "C5": dict(model="ellipse", n_scene=50000, n_model=50000, H=65536, max_iter=10),
        "tiny": dict(model="ellipse", n_scene=500, n_model=2000, H=64, max_iter=10),
        # one rank's share of the headline frame under strong scaling (N = 8, 4, 2): batch-shape experiments
        "shard2k": dict(model="ellipse", n_scene=10000, n_model=10000, H=2048, max_iter=10),
        "shard4k": dict(model="ellipse", n_scene=10000, n_model=10000, H=4096, max_iter=10),
        "shard8k": dict(model="ellipse", n_scene=10000, n_model=10000, H=8192, max_iter=10),
    }
    return table[name]


# ------------------------------------------------------------------------------------------------------------------
# hand-state search fixtures (K1): a two-finger gripper in the hand-base frame, SURVEY.md 8(d) sizes
# ------------------------------------------------------------------------------------------------------------------
def _rot_x(th):
    c, s = np.cos(th), np.sin(th)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1.0]])


def make_finger_cloud(n, seed=0, size=(0.02, 0.012, 0.06)):
    """Surface samples of a finger link: a box, x in [-sx/2,sx/2], y in [-sy/2,sy/2] (+y = inner face), z in [-sz,0]
    (the tip is at min z, the joint axis is local x through the origin), outward unit normals."""
    rng = np.random.default_rng(seed)
    pts, nrm = _cuboid(rng, n, *size)
    pts[:, 2] -= size[2] / 2
    return pts.astype(np.float32), nrm.astype(np.float32)


def make_hand_case(seed=0, n_finger=300, n_hand=3000, theta_true_deg=15.0, palm_side=False, right_side=False, bad_normals=True,
                   shifted_lookup=False):
    """One finger-link search problem.  Returns a dict with the clouds (finger, scene_hand = what the kd-tree indexes,
    scene_lookup = the cloud the reference reads the neighbour's normal from, scene_noswivel) and the scalar fields of
    hop_finger_params (everything except the FingerProperty histogram, which the host mirror derives from the cloud)."""
    rng = np.random.default_rng(seed)
    f_xyz, f_nrm = make_finger_cloud(n_finger, seed=seed + 1)
    sgn = 1.0 if not right_side else -1.0            # left finger sits at -y and closes towards +y; the right one is mirrored
    m2hb = np.eye(4)
    if right_side:
        m2hb[:3, :3] = np.diag([-1.0, -1.0, 1.0])    # rotate 180 deg about z: local +y (inner) faces hand-base -y
    m2hb[:3, 3] = [-0.15, -0.04 * sgn, 0.02]
    th = np.deg2rad(theta_true_deg)
    T_true = m2hb @ _rot_x(th)
    # scene: the true finger (dense, noisy), a palm plate, the grasped object between the fingers, clutter
    n_f, n_p, n_o = int(0.45 * n_hand), int(0.2 * n_hand), int(0.25 * n_hand)
    n_c = n_hand - n_f - n_p - n_o
    fp, fn = _cuboid(rng, n_f, 0.02, 0.012, 0.06)
    fp[:, 2] -= 0.03
    fp = fp + fn * rng.normal(0, 0.0003, (n_f, 1))
    Pf = fp @ T_true[:3, :3].T + T_true[:3, 3]
    Nf = fn @ T_true[:3, :3].T
    pp, pn = _cuboid(rng, n_p, 0.06, 0.06, 0.01)
    Pp = pp + [-0.15, 0.0, 0.03]
    po, no = _ellipsoid(rng, n_o, 0.02, 0.018, 0.015)
    Po = po + [-0.15, 0.0, -0.03]
    Pc = rng.uniform([-0.3, -0.03, -0.06], [-0.05, 0.03, 0.03], (n_c, 3))
    Nc = rng.normal(size=(n_c, 3))
    Nc /= np.linalg.norm(Nc, axis=1, keepdims=True)
    xyz = np.concatenate([Pf, Pp, Po, Pc]).astype(np.float32)
    nrm = np.concatenate([Nf, pn, no, Nc]).astype(np.float32)
    if bad_normals:                                   # the reference branches on all-zero and non-finite normals
        k = rng.choice(n_f, size=max(n_f // 20, 2), replace=False)
        nrm[k[: len(k) // 2]] = 0.0
        nrm[k[len(k) // 2:], rng.integers(0, 3)] = np.nan
    perm = rng.permutation(len(xyz))
    xyz, nrm = xyz[perm], nrm[perm]
    # scene_hand_region (lookup) = the unfiltered cloud; scene_hand (kd-tree) = a subset with shifted indices
    # (shifted_lookup: the noise filters dropped points anywhere, so the reference's index reuse reads unrelated normals;
    #  otherwise only trailing points were dropped and the indices still line up)
    keep = rng.random(len(xyz)) > 0.03 if shifted_lookup else np.arange(len(xyz)) < int(0.97 * len(xyz))
    scene_xyz, scene_nrm = xyz[keep], nrm[keep]
    sw = (scene_xyz[:, 0] > -0.25) & (scene_xyz[:, 0] < -0.1)  # Hand.cpp:312-318 pass-through
    f_min, f_max = f_xyz.min(0), f_xyz.max(0)
    if palm_side:
        tip1 = [f_min[0], f_max[1], f_min[2]]         # of the distal link, in ITS frame (finger_out_property)
        tip2 = [f_min[0], f_max[1], f_min[2]]
    else:
        tip1 = [f_min[0], f_max[1], f_min[2]]
        tip2 = [f_min[0], f_max[1], f_max[2]]
    out2parent = np.eye(4)
    out2parent[:3, 3] = [0, 0, -0.06]                 # distal link hangs at the proximal link's tip
    scalars = dict(model2handbase=m2hb.astype(np.float32), finger_out2parent=out2parent.astype(np.float32),
                   tip1_local=np.array(tip1, np.float32), tip2_local=np.array(tip2, np.float32),
                   pair_tip1_y=float(0.02 * sgn), pair_tip2_y=float(0.035 * sgn), palm_side=int(palm_side), right_side=int(right_side),
                   gripper_min_dist=0.03, dist_thres=0.005, normal_angle_deg=60.0, check_normal=1, num_division=10,
                   max_outter_pts=300, outter_pt_dist=0.002, outter_pt_dist_weight=1.0)
    return dict(finger_xyz=f_xyz, finger_nrm=f_nrm, scene_xyz=scene_xyz, scene_nrm=scene_nrm, lookup_xyz=xyz, lookup_nrm=nrm,
                noswivel_xyz=scene_xyz[sw], noswivel_nrm=scene_nrm[sw], theta_true=th, scalars=scalars)


def hand_variants():
    """name -> (make_hand_case kwargs, scalar overrides): together they reach every branch of objFuncPSO."""
    return {
        "left": (dict(), dict()),
        "right": (dict(right_side=True), dict()),
        "palm": (dict(palm_side=True), dict()),
        "nogap": (dict(), dict(gripper_min_dist=-1.0, dist_thres=0.0005)),   # no gap penalty, 0.5 mm gate: no match -> 100 - theta
        "exp": (dict(), dict(outter_pt_dist=1e-4, max_outter_pts=10 ** 6)),  # the exp(avg * 1000) branch
        "nonormal": (dict(), dict(check_normal=0)),
        "shifted": (dict(shifted_lookup=True), dict()),                      # index reuse reads unrelated normals
        "tight": (dict(n_hand=1200, n_finger=150), dict(dist_thres=0.002, normal_angle_deg=30.0)),
    }


def hand_problem(variant, seed=5):
    kw, over = hand_variants()[variant]
    case = make_hand_case(seed=seed, **kw)
    case["scalars"].update(over)
    return case


# ------------------------------------------------------------------------------------------------------------------
# triangle meshes + grasp scenes for the physics pruning step (PoseEstimator::rejectByCollisionOrNonTouching)
# ------------------------------------------------------------------------------------------------------------------
def _orient_outward(V, F):
    """flip faces whose normal points towards the (star-shaped) mesh's centroid"""
    c = V.mean(0)
    a, b, d = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    n = np.cross(b - a, d - a)
    flip = np.einsum("ij,ij->i", n, (a + b + d) / 3 - c) < 0
    F = F.copy()
    F[flip] = F[flip][:, [0, 2, 1]]
    return F


def _icosphere(level):
    t = (1 + 5 ** 0.5) / 2
    V = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1),
         (-t, 0, -1), (-t, 0, 1)]
    F = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    V = [np.array(v, float) / np.linalg.norm(v) for v in V]
    for _ in range(level):
        cache, F2 = {}, []

        def mid(i, j):
            key = (min(i, j), max(i, j))
            if key not in cache:
                m = V[i] + V[j]
                V.append(m / np.linalg.norm(m))
                cache[key] = len(V) - 1
            return cache[key]
        for a, b, c in F:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            F2 += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        F = F2
    return np.array(V), np.array(F, np.int32)


def _box_mesh(size, div=1, offset=(0, 0, 0)):
    """closed box, each face a div x div grid of quads split in two; vertices shared along edges"""
    h = np.asarray(size, float) / 2
    key, V, F = {}, [], []

    def vid(p):
        k = tuple(np.round(p / (h / div)).astype(int))
        if k not in key:
            key[k] = len(V)
            V.append(p)
        return key[k]
    for ax in range(3):
        u, w = (ax + 1) % 3, (ax + 2) % 3
        for sg in (1.0, -1.0):
            for i in range(div):
                for j in range(div):
                    q = []
                    for di, dj in ((0, 0), (1, 0), (1, 1), (0, 1)):
                        p = np.zeros(3)
                        p[ax] = sg * h[ax]
                        p[u] = -h[u] + 2 * h[u] * (i + di) / div
                        p[w] = -h[w] + 2 * h[w] * (j + dj) / div
                        q.append(vid(p))
                    F += [(q[0], q[1], q[2]), (q[0], q[2], q[3])]
    V, F = np.array(V), np.array(F, np.int32)
    return V + np.asarray(offset, float), _orient_outward(V, F)


def _lathe(profile, nseg):
    """revolve an (r, z) polyline that starts and ends on the axis (r = 0) about z"""
    prof = np.asarray(profile, float)
    V, ring = [], []
    for r, z in prof:
        if r == 0:
            V.append((0, 0, z))
            ring.append([len(V) - 1] * nseg)
        else:
            ids = []
            for s in range(nseg):
                th = 2 * np.pi * s / nseg
                V.append((r * np.cos(th), r * np.sin(th), z))
                ids.append(len(V) - 1)
            ring.append(ids)
    F = []
    for k in range(len(prof) - 1):
        a, b = ring[k], ring[k + 1]
        for s in range(nseg):
            t = (s + 1) % nseg
            if a[s] != a[t]:
                F.append((a[s], a[t], b[t]))
            if b[s] != b[t]:
                F.append((a[s], b[t], b[s]))
    V, F = np.array(V), np.array(F, np.int32)
    # a lathe surface is star-shaped about its axis per z-slab, not about the centroid: orient by the radial/axial direction
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    n = np.cross(b - a, c - a)
    m = (a + b + c) / 3
    radial = np.stack([m[:, 0], m[:, 1], np.zeros(len(m))], 1)
    side = np.abs(n[:, 2]) < 1e-12 * np.linalg.norm(n, axis=1).max()
    out = np.where(side, np.einsum("ij,ij->i", n, radial), 0.0)
    flip = out < 0
    # caps / flange rings: decide by a ray test against the analytic profile (outward = away from the solid)
    for i in np.where(~side)[0]:
        r, z = np.hypot(m[i, 0], m[i, 1]), m[i, 2]
        up = _lathe_inside(prof, r, z + 1e-6)
        flip[i] = (n[i, 2] > 0) == up
    F = F.copy()
    F[flip] = F[flip][:, [0, 2, 1]]
    return V, F


def _lathe_inside(prof, r, z):
    """point-in-polygon of (r, z) against the closed profile (axis closes it)"""
    poly = np.concatenate([prof, prof[:1]])
    inside = False
    for (r0, z0), (r1, z1) in zip(poly[:-1], poly[1:]):
        if (z0 > z) != (z1 > z):
            rc = r0 + (z - z0) * (r1 - r0) / (z1 - z0)
            if rc > r:
                inside = not inside
    return inside


def make_mesh(name, level=2):
    """closed, outward-oriented triangle mesh of a synthetic object (float32 vertices, int32 faces); `level` refines it"""
    if name == "ellipse":
        V, F = _icosphere(level)
        V = V * [0.045, 0.030, 0.020]
        F = _orient_outward(V, F)
    elif name == "cuboid":
        V, F = _box_mesh((0.08, 0.05, 0.03), div=max(1, level))
    elif name == "cylinder":
        V, F = _lathe([(0, -0.045), (0.025, -0.045), (0.025, 0.045), (0, 0.045)], 8 * max(1, level))
    elif name == "tless":
        V, F = _lathe([(0, -0.03), (0.045, -0.03), (0.045, -0.02), (0.03, -0.02), (0.03, 0.02), (0, 0.02)], 8 * max(1, level))
    else:
        raise KeyError(name)
    return np.ascontiguousarray(V, np.float32), np.ascontiguousarray(F, np.int32)


def make_collision_case(name="ellipse", H=256, seed=0, n_model=500, n_finger=300, n_scene=2500, n_hand=1500, mesh_level=2,
                        rot_sigma_deg=8.0, trans_sigma=0.008, disabled=()):
    """A grasp in the hand-base frame: the object between two two-link fingers, a camera looking at it, H pose hypotheses (in the
    camera frame, like PoseHypo::_pose) scattered around the true pose widely enough to hit every reject branch.
    Returns a dict with everything hop_reject_by_collision / the oracle take.  Finger order: finger_1_1, finger_1_2, finger_2_1,
    finger_2_2 (proximal, distal of finger 1; proximal, distal of finger 2)."""
    rng = np.random.default_rng(seed)
    V, F = make_mesh(name, mesh_level)
    m_xyz, _ = make_model(name, n_model, seed=seed + 11)
    ext = V.max(0) - V.min(0)
    smallest, diam = float(ext.min()), float(np.linalg.norm(ext))
    # object pose in the hand-base frame: centred between the fingers, which close along y
    R = random_rotation(rng)
    obj_in_hb = np.eye(4)
    obj_in_hb[:3, :3] = R
    obj_in_hb[:3, 3] = [-0.17, 0.0, -0.02]
    Vh = V @ R.T + obj_in_hb[:3, 3]
    half_y = float(np.abs(Vh[:, 1]).max())
    # camera: looks along +z from 0.35 m
    cam_in_hb = np.eye(4)
    cam_in_hb[:3, :3] = _rot_x(np.deg2rad(200.0))[:3, :3] @ _rot_from_rotvec(np.array([0.0, 0.25, 0.1]))
    cam_in_hb[:3, 3] = [-0.15, 0.05, 0.33]
    cam2hb = cam_in_hb                                  # maps camera-frame points into the hand base (handbase_in_cam^-1)
    gt_cam = np.linalg.inv(cam2hb) @ obj_in_hb
    # fingers: boxes (link meshes) whose inner faces touch the object from -y and +y
    size = (0.02, 0.012, 0.05)
    fingers_V, fingers_F, fingers_pts = [], [], []
    for side in (-1.0, 1.0):
        for link in range(2):
            off = np.array([-0.17 + (0.03 if link == 0 else -0.025), side * (half_y + size[1] / 2 + 0.0005), -0.02])
            fv, ff = _box_mesh(size, div=3, offset=off)
            fp, _ = _cuboid(rng, n_finger, *size)
            fingers_V.append(fv.astype(np.float32))
            fingers_F.append(ff)
            fingers_pts.append((fp + off).astype(np.float32))
    status = np.array([0 if k in disabled else 1 for k in range(4)], np.int32)
    hand_xyz = np.concatenate(fingers_pts + [(_cuboid(rng, n_hand // 3, 0.08, 0.1, 0.02)[0] + [-0.09, 0, -0.02]).astype(np.float32)])
    hand_xyz = hand_xyz[rng.permutation(len(hand_xyz))[:n_hand]].astype(np.float32)
    # scene without the hand: visible part of the object + table clutter, in the hand-base frame, 5 mm voxel sampled
    sp, sn = make_model(name, 6 * n_scene, seed=seed + 12)
    sp = sp @ R.T + obj_in_hb[:3, 3]
    view = cam_in_hb[:3, 3] - sp
    vis = np.einsum("ij,ij->i", sn @ R.T, view) > 0
    sp = sp[vis] + rng.normal(0, 0.0004, (int(vis.sum()), 3))
    clutter = rng.uniform([-0.3, -0.15, -0.12], [0.0, 0.15, -0.10], (n_scene, 3))
    scene = np.concatenate([sp, clutter])
    vox = np.unique(np.floor(scene / 0.005).astype(np.int64), axis=0, return_index=True)[1]
    scene = scene[np.sort(vox)][:n_scene].astype(np.float32)
    poses = make_hypotheses(gt_cam, H, seed=seed + 13, rot_sigma_deg=rot_sigma_deg, trans_sigma=trans_sigma, random_frac=0.05)
    params = dict(cam2handbase=cam2hb.astype(np.float32), model_center=m_xyz.mean(0).astype(np.float32),
                  ob_diameter=diam, collision_dist=min(-smallest * 0.4, -0.007), inside_ob_dist=min(-smallest / 5, -0.01),
                  non_touch_dist=0.01, collision_finger_dist=-0.012, collision_finger_volume_ratio=0.25, finger_status=status)
    # (config_autodataset.yaml:128-131: collision_thres 0.4, non_touch_dist 0.01, collision_finger_dist 0.012, volume ratio 0.25)
    return dict(obj_V=V, obj_F=F, finger_V=fingers_V, finger_F=fingers_F, finger_pts=fingers_pts, hand_xyz=hand_xyz, scene_xyz=scene,
                model_xyz=m_xyz, poses=poses.astype(np.float32), gt=gt_cam.astype(np.float32), params=params)


def make_render_case(name="ellipse", H=64, seed=0, width=640, height=480, noise=0.001, dropout=0.03, mesh_level=2, render=None):
    """A frame for PoseEstimator::rejectByRender: the grasp of make_collision_case seen by the camera.  `render(p_kw, hand_V, hand_F,
    obj_V, obj_F, pose) -> depth` produces the "real" depth image from the true pose (the tests pass the oracle's renderer); noise,
    a background plane at 0.8 m where nothing is hit, and dropouts (0 = invalid, like the sensor) are added on top.
    Returns dict(obj_V, obj_F, hand_V, hand_F (camera frame, concatenated), poses (camera frame), gt, depth_m, cam = dict(fx, fy, cx, cy, width, height))."""
    case = make_collision_case(name, H=H, seed=seed, mesh_level=mesh_level)
    rng = np.random.default_rng(seed + 101)
    # a camera that looks at the grasp (the collision case's camera only has to define a frame): 0.35 m away, slightly off-axis
    old_cam2hb = case["params"]["cam2handbase"].astype(np.float64)
    target = np.array([-0.17, 0.0, -0.02])
    eye = target + np.array([0.03, 0.07, 0.34])
    zc = (target - eye) / np.linalg.norm(target - eye)
    xc = np.cross([0.0, 1.0, 0.0], zc); xc /= np.linalg.norm(xc)
    yc = np.cross(zc, xc)
    cam2hb = np.eye(4)
    cam2hb[:3, :3] = np.stack([xc, yc, zc], 1)
    cam2hb[:3, 3] = eye
    hb2cam = np.linalg.inv(cam2hb)
    regauge = hb2cam @ old_cam2hb                       # the hypotheses and the true pose, re-expressed in the new camera frame
    case["poses"] = np.einsum("ij,hjk->hik", regauge, case["poses"].astype(np.float64)).astype(np.float32)
    case["gt"] = (regauge @ case["gt"].astype(np.float64)).astype(np.float32)
    case["params"] = dict(case["params"], cam2handbase=cam2hb.astype(np.float32))
    hV, hF, off = [], [], 0
    for v, f in zip(case["finger_V"], case["finger_F"]):
        hV.append((v.astype(np.float64) @ hb2cam[:3, :3].T + hb2cam[:3, 3]).astype(np.float32))
        hF.append(f + off)
        off += len(v)
    hand_V, hand_F = np.concatenate(hV), np.concatenate(hF).astype(np.int32)
    s = width / 640.0
    cam = dict(fx=616.596 * s, fy=616.596 * s, cx=307.628 * s, cy=239.687 * s, width=width, height=height)
    out = dict(obj_V=case["obj_V"], obj_F=case["obj_F"], hand_V=hand_V, hand_F=hand_F, poses=case["poses"], gt=case["gt"], cam=cam,
               collision=case)            # the same grasp as a make_collision_case dict in the new camera's gauge
    if render is not None:
        d = render(cam, hand_V, hand_F, case["obj_V"], case["obj_F"], case["gt"]).astype(np.float32)
        hit = d < 1.999
        d = np.where(hit, d + rng.normal(0, noise, d.shape), 0.8 + rng.normal(0, noise, d.shape)).astype(np.float32)
        d[rng.random(d.shape) < dropout] = 0.0
        out["depth_m"] = d
    return out


# ------------------------------------------------------------------------------------------------------------------
# hand-point removal fixtures (HandT42::removeSurroundingPointsAndAssignProbability)
# ------------------------------------------------------------------------------------------------------------------
HAND_LINKS = ("base", "finger_1_1", "finger_1_2", "finger_2_1", "finger_2_2", "swivel_1", "swivel_2")   # std::map order
HAND_LINK_KIND = (2, 1, 0, 1, 0, 2, 2)


def make_hand_removal_case(seed=0, n_scene=6000, n_link=400):
    """A cropped scene in the camera frame (hand + grasped object + clutter, with normals) and the seven link clouds of a T42-like
    hand in the hand-base frame.  Returns dict(scene_xyz, scene_nrm, links (list in HAND_LINKS order), kinds, handbase_in_cam,
    finger_1_2_in_handbase, finger_2_2_in_handbase, min_z, near_hand_dist)."""
    rng = np.random.default_rng(seed)
    boxes = {"base": ((0.08, 0.10, 0.03), (-0.06, 0.0, 0.0)), "swivel_1": ((0.03, 0.03, 0.03), (-0.10, -0.04, 0.0)),
             "swivel_2": ((0.03, 0.03, 0.03), (-0.10, 0.04, 0.0)), "finger_1_1": ((0.05, 0.012, 0.02), (-0.14, -0.045, 0.0)),
             "finger_1_2": ((0.05, 0.012, 0.02), (-0.19, -0.04, 0.0)), "finger_2_1": ((0.05, 0.012, 0.02), (-0.14, 0.045, 0.0)),
             "finger_2_2": ((0.05, 0.012, 0.02), (-0.19, 0.04, 0.0))}
    links = []
    for name in HAND_LINKS:
        size, off = boxes[name]
        p, _ = _cuboid(rng, n_link, *size)
        links.append((p + off).astype(np.float32))
    # distal link frames: origin at the link's centre, +y towards the other finger (so y < 0 is the outer side), z along the hand's x
    def frame(off, inward):
        T = np.eye(4)
        T[:3, 0] = [0, 0, 1.0]
        T[:3, 1] = [0, inward, 0]
        T[:3, 2] = np.cross(T[:3, 0], T[:3, 1])
        T[:3, 3] = off
        return T
    f12 = frame(boxes["finger_1_2"][1], 1.0)
    f22 = frame(boxes["finger_2_2"][1], -1.0)
    n_h, n_o = n_scene // 2, n_scene // 4
    pick = rng.integers(0, len(HAND_LINKS), n_h)
    hand_pts = np.stack([links[k][rng.integers(0, n_link)] for k in pick]) + rng.normal(0, 0.0015, (n_h, 3))
    obj, _ = _ellipsoid(rng, n_o, 0.03, 0.025, 0.02)
    obj = obj + [-0.19, 0.0, 0.0]
    clutter = rng.uniform([-0.3, -0.12, -0.08], [0.0, 0.12, 0.08], (n_scene - n_h - n_o, 3))
    hb = np.concatenate([hand_pts, obj, clutter])
    hb = hb[rng.permutation(len(hb))]
    nrm = rng.normal(size=hb.shape)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    hic = np.eye(4)
    hic[:3, :3] = random_rotation(rng)
    hic[:3, 3] = [0.1, -0.05, 0.45]
    scene = hb @ hic[:3, :3].T + hic[:3, 3]
    return dict(scene_xyz=scene.astype(np.float32), scene_nrm=(nrm @ hic[:3, :3].T).astype(np.float32), links=links, kinds=np.array(HAND_LINK_KIND, np.int32),
                handbase_in_cam=hic.astype(np.float32), finger_1_2_in_handbase=f12.astype(np.float32), finger_2_2_in_handbase=f22.astype(np.float32),
                min_z=-0.02, near_hand_dist=0.003)
