"""ctypes access to the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
liboracle.so = plain-C restatement (hop_oracle*.c); _ref/libhop_ref.so = pieces compiled from the reference tree.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(verbose=False):
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        L = _lib
        L.hop_oracle_nn.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, _i32p, _f32p]
        L.hop_oracle_transform_cloud.argtypes = [_f32p, _f32p, _f32p, C.c_int, _f32p, _f32p]
        L.hop_oracle_compute_lcp.restype = C.c_float
        L.hop_oracle_compute_lcp.argtypes = [_f32p, _f32p, C.c_int, C.c_void_p, _f32p, _f32p, C.c_int, C.c_float,
                                             C.c_float, C.c_int, C.c_int, C.c_int]
        L.hop_oracle_warp6d.argtypes = [_f32p, _f32p]
        L.hop_oracle_lm_point_to_plane.restype = C.c_int
        L.hop_oracle_lm_point_to_plane.argtypes = [_f32p, _f32p, _f32p, C.c_int, _f32p, C.POINTER(C.c_int)]
        L.hop_oracle_set_lm_backend.argtypes = [C.c_void_p]
        L.hop_oracle_run_icp.restype = C.c_int
        L.hop_oracle_run_icp.argtypes = [_f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int, C.c_float,
                                         C.c_float, C.c_double, C.POINTER(C.c_int)]
        L.hop_oracle_refine_by_icp.argtypes = [_f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int, C.c_int,
                                               C.c_float, C.c_float, C.c_double, C.c_int, _i32p, _i32p]
        L.hop_oracle_select_best.restype = C.c_int
        L.hop_oracle_select_best.argtypes = [_f32p, _f32p, C.c_int, C.c_void_p, _f32p, _f32p, C.c_int, _f32p, C.c_int,
                                             C.c_float, C.c_float, C.c_int, _f32p]
        L.hop_oracle_verify_quads.restype = C.c_int
        L.hop_oracle_verify_quads.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, _i32p, _i32p, C.c_int, _f32p, _f32p, C.c_float,
                                              _f32p, _f32p, _i32p, C.c_int]
        L.hop_oracle_num_threads.restype = C.c_int
        _f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.hop_oracle_obj_func_pso.restype = C.c_double
        L.hop_oracle_obj_func_pso.argtypes = [C.c_double, C.c_void_p, _f32p, _f32p, C.c_int, _f32p, C.c_int, _f32p, _f32p, C.c_int, C.c_void_p]
        L.hop_oracle_hand_overlap.argtypes = [C.c_void_p, _f32p, _f32p, C.c_int, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f64p, C.c_int,
                                              _f64p, C.c_void_p, C.c_int]
        L.hop_oracle_finger_property.argtypes = [_f32p, C.c_int, C.c_int, C.c_void_p, _f32p]
        L.hop_oracle_hand_tf_self.argtypes = [C.c_double, _f32p]
        L.hop_oracle_hand_inverse.argtypes = [_f32p, _f32p]
    return _lib


def ref():
    """oracle/_ref/libhop_ref.so or None when it was never built (no /root/reference at build time)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libhop_ref.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        _ref.hop_ref_lm_point_to_plane.restype = C.c_int
        _ref.hop_ref_lm_point_to_plane.argtypes = [_f32p, _f32p, _f32p, C.c_int, _f32p, C.POINTER(C.c_int)]
        if hasattr(_ref, "hop_ref_cluster_poses"):
            _ref.hop_ref_cluster_poses.restype = C.c_int
            _ref.hop_ref_cluster_poses.argtypes = [_f32p, _f32p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                                   np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS"), _i32p]
            _ref.hop_ref_euler_zyx.argtypes = [_f32p, _f32p]
        if hasattr(_ref, "hop_ref_s4pcs_run"):
            _ref.hop_ref_s4pcs_run.restype = C.c_int
            _ref.hop_ref_s4pcs_run.argtypes = [_f32p, _f32p, C.c_void_p, C.c_int, _f32p, _f32p, C.c_int, _i32p, C.c_int,
                                               C.POINTER(S4pcsOptions)]
            _ref.hop_ref_s4pcs_sizes.argtypes = [_i32p]
            _ref.hop_ref_s4pcs_get.argtypes = [_f32p, _f32p, _i32p, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p]
            _ref.hop_ref_ppf_pairs.argtypes = [_f32p, _f32p, C.c_int, _i32p]
            _ref.hop_ref_s4pcs_get_quads.argtypes = [_i32p, _f32p, _f32p]
            _ref.hop_ref_s4pcs_num_pairs.restype = C.c_int
            _ref.hop_ref_s4pcs_get_trials.argtypes = [_i32p, _f32p, _i32p]
            _ref.hop_ref_compute_ppf.argtypes = [_f32p, _f32p, _f32p, _f32p, _i32p]
    return _ref


class S4pcsOptions(C.Structure):
    """the options PoseEstimator::runSuper4pcs sets (PoseEstimator.cpp:66-73), defaults = config_autodataset.yaml:133-140"""
    _fields_ = [("sample_size", C.c_int32), ("overlap", C.c_float), ("delta", C.c_float), ("dispersion", C.c_float),
                ("success_quadrilaterals", C.c_int32), ("max_normal_difference", C.c_float), ("max_color_distance", C.c_float),
                ("max_trials", C.c_int32), ("nthreads", C.c_int32)]

    def __init__(self, **kw):
        super().__init__(sample_size=100, overlap=0.2, delta=0.003, dispersion=0.5, success_quadrilaterals=10,
                         max_normal_difference=-1.0, max_color_distance=-1.0, max_trials=0, nthreads=1)
        for k, v in kw.items():
            setattr(self, k, v)


def ref_ppf_keys(xyz, nrm):
    """unique PPF keys of all point pairs of a cloud, by the reference's own gr::computePPF (compiled in oracle/_ref)."""
    R = ref()
    xyz, nrm = _c(xyz), _c(nrm)
    n = len(xyz)
    keys = np.zeros((n * (n - 1) // 2, 4), np.int32)
    R.hop_ref_ppf_pairs(xyz, nrm, n, keys)
    return np.unique(keys, axis=0)


def ref_super4pcs(P, Pn, Pprob, Q, Qn, ppf_keys, **opts):
    """The reference's compiled Super4PCS matcher (oracle/_ref).  Returns a dict: poses (H,4,4), lcp (H,), trials (T,9)
    [base0..3, ok, quad_begin, quad_end, hyp_begin, hyp_end], quads (M,4), centred sampled clouds and centroids."""
    R = ref()
    o = S4pcsOptions(**opts)
    P, Pn, Q, Qn = _c(P), _c(Pn), _c(Q), _c(Qn)
    keys = np.ascontiguousarray(ppf_keys, np.int32).reshape(-1, 4)
    prob = None if Pprob is None else _c(Pprob)
    H = R.hop_ref_s4pcs_run(P, Pn, None if prob is None else prob.ctypes.data_as(C.c_void_p), len(P), Q, Qn, len(Q), keys, len(keys), C.byref(o))
    sizes = np.zeros(5, np.int32)
    R.hop_ref_s4pcs_sizes(sizes)
    H, T, M, nP, nQ = [int(x) for x in sizes]
    poses = np.zeros((max(H, 1), 16), np.float32); lcp = np.zeros(max(H, 1), np.float32)
    trials = np.zeros((max(T, 1), 9), np.int32); quads = np.zeros((max(M, 1), 4), np.int32)
    Pc = np.zeros((nP, 3), np.float32); Pcn = np.zeros((nP, 3), np.float32)
    Qc = np.zeros((nQ, 3), np.float32); Qcn = np.zeros((nQ, 3), np.float32)
    cen = np.zeros(6, np.float32); misc = np.zeros(4, np.float32)
    R.hop_ref_s4pcs_get(poses, lcp, trials, quads, Pc, Pcn, Qc, Qcn, cen, misc)
    q_ok = np.zeros(max(M, 1), np.int32); q_rms = np.zeros(max(M, 1), np.float32); q_lcp = np.zeros(max(M, 1), np.float32)
    R.hop_ref_s4pcs_get_quads(q_ok, q_rms, q_lcp)
    npairs = R.hop_ref_s4pcs_num_pairs()
    t_i = np.zeros((max(T, 1), 9), np.int32); t_f = np.zeros((max(T, 1), 4), np.float32); pairs = np.zeros((max(npairs, 1), 2), np.int32)
    R.hop_ref_s4pcs_get_trials(t_i, t_f, pairs)
    extra = dict(base_ok=t_i[:T, 0].copy(), bases_all=t_i[:T, 1:5].copy(), pair_ranges=t_i[:T, 5:9].copy(), inv=t_f[:T, :2].copy(),
                 dist=t_f[:T, 2:].copy(), pairs=pairs[:npairs])
    return dict(**extra, quad_ok=q_ok[:M], quad_rms=q_rms[:M], quad_lcp=q_lcp[:M], poses=colmajor_to_poses(poses[:H]), lcp=lcp[:H], trials=trials[:T], quads=quads[:M], Pc=Pc, Pn=Pcn, Qc=Qc, Qn=Qcn,
                centroid_P=cen[:3].copy(), centroid_Q=cen[3:].copy(), diameter=float(misc[0]), delta=o.delta)


def verify_quads(Pc, Qc, bases, quads, quad_trial, centroid_P, centroid_Q, delta, nthreads=1):
    """C restatement of TryCongruentSet + ComputeRigidTransformation + Verify (hop_oracle_verify_quads)."""
    Pc, Qc = _c(Pc), _c(Qc)
    bases = np.ascontiguousarray(bases, np.int32).reshape(-1, 4)
    quads = np.ascontiguousarray(quads, np.int32).reshape(-1, 4)
    qt = np.ascontiguousarray(quad_trial, np.int32)
    M = len(quads)
    poses = np.zeros((max(M, 1), 16), np.float32); lcp = np.zeros(max(M, 1), np.float32); valid = np.zeros(max(M, 1), np.int32)
    n = lib().hop_oracle_verify_quads(Pc, len(Pc), Qc, len(Qc), bases, quads, qt, M, _c(centroid_P), _c(centroid_Q), delta, poses, lcp,
                                      valid, nthreads)
    return colmajor_to_poses(poses[:M]), lcp[:M], valid[:M], n


def quad_trial_of(trials, n_quads):
    """per-quadrilateral trial index from the (T,9) trial table of ref_super4pcs"""
    qt = np.zeros(n_quads, np.int32)
    for i, t in enumerate(trials):
        qt[t[5]:t[6]] = i
    return qt


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def poses_to_colmajor(poses):
    """(H,4,4) row-major numpy matrices -> (H,16) column-major float32 (Eigen::Matrix4f::data())."""
    p = np.asarray(poses, dtype=np.float32).reshape(-1, 4, 4)
    return np.ascontiguousarray(p.transpose(0, 2, 1).reshape(-1, 16))


def colmajor_to_poses(flat):
    return np.asarray(flat, dtype=np.float32).reshape(-1, 4, 4).transpose(0, 2, 1).copy()


def nn(pts, q, use_kdtree=True):
    pts, q = _c(pts), _c(q)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float32)
    lib().hop_oracle_nn(pts, len(pts), q, len(q), int(use_kdtree), idx, d2)
    return idx, d2


def transform_cloud(T, xyz, nrm):
    xyz, nrm = _c(xyz), _c(nrm)
    oxyz, onrm = np.empty_like(xyz), np.empty_like(nrm)
    lib().hop_oracle_transform_cloud(poses_to_colmajor(T)[0], xyz, nrm, len(xyz), oxyz, onrm)
    return oxyz, onrm


def compute_lcp(s_xyz, s_nrm, m_xyz, m_nrm, dist, angle, weights=None, use_normal=True, use_dot=True, use_recip=True):
    s_xyz, s_nrm, m_xyz, m_nrm = _c(s_xyz), _c(s_nrm), _c(m_xyz), _c(m_nrm)
    w = None if weights is None else _c(weights).ctypes.data_as(C.c_void_p)
    return float(lib().hop_oracle_compute_lcp(s_xyz, s_nrm, len(s_xyz), w, m_xyz, m_nrm, len(m_xyz), dist, angle,
                                              int(use_normal), int(use_dot), int(use_recip)))


def lm_point_to_plane(src, tgt, nrm, backend="c"):
    src, tgt, nrm = _c(src), _c(tgt), _c(nrm)
    x = np.zeros(6, np.float32)
    nfev = C.c_int(0)
    if backend == "c":
        info = lib().hop_oracle_lm_point_to_plane(src, tgt, nrm, len(src), x, C.byref(nfev))
    else:
        info = ref().hop_ref_lm_point_to_plane(src, tgt, nrm, len(src), x, C.byref(nfev))
    return x, info, nfev.value


def use_ref_lm(enable=True):
    """Route the LM step of hop_oracle_run_icp through the reference tree's own Eigen LM (oracle/_ref)."""
    if enable:
        r = ref()
        if r is None:
            raise RuntimeError("oracle/_ref not built")
        lib().hop_oracle_set_lm_backend(C.cast(r.hop_ref_lm_point_to_plane, C.c_void_p))
    else:
        lib().hop_oracle_set_lm_backend(None)


def warp6d(x):
    T = np.empty(16, np.float32)
    lib().hop_oracle_warp6d(_c(x), T)
    return colmajor_to_poses(T)[0]


def warp_y13(x):
    """y(x) = [vec(R(x)) - vec(I); t(x); 1] of csrc/lm_replay.cuh, from the oracle's WarpPointRigid6D (float64 array of 13)."""
    T = warp6d(x).astype(np.float64)
    return np.concatenate([(T[:3, :3] - np.eye(3)).reshape(-1), T[:3, 3], [1.0]])


def run_icp(src_xyz, src_nrm, tgt_xyz, tgt_nrm, max_iter=10, angle=45.0, dist=0.01, abs_mse_eps=1e-6):
    src_xyz, src_nrm, tgt_xyz, tgt_nrm = _c(src_xyz), _c(src_nrm), _c(tgt_xyz), _c(tgt_nrm)
    T = np.empty(16, np.float32)
    conv = C.c_int(0)
    it = lib().hop_oracle_run_icp(src_xyz, src_nrm, len(src_xyz), tgt_xyz, tgt_nrm, len(tgt_xyz), T, max_iter, angle,
                                  dist, abs_mse_eps, C.byref(conv))
    return colmajor_to_poses(T)[0], it, bool(conv.value)


def refine_by_icp(s_xyz, s_nrm, m_xyz, m_nrm, poses, max_iter=10, angle=45.0, dist=0.01, abs_mse_eps=1e-6, nthreads=0):
    """poses: (H,4,4) model2scene. Returns refined (H,4,4), iterations (H,), converged (H,)."""
    s_xyz, s_nrm, m_xyz, m_nrm = _c(s_xyz), _c(s_nrm), _c(m_xyz), _c(m_nrm)
    flat = poses_to_colmajor(poses)
    H = len(flat)
    iters = np.zeros(H, np.int32)
    conv = np.zeros(H, np.int32)
    lib().hop_oracle_refine_by_icp(s_xyz, s_nrm, len(s_xyz), m_xyz, m_nrm, len(m_xyz), flat, H, max_iter, angle, dist,
                                   abs_mse_eps, nthreads, iters, conv)
    return colmajor_to_poses(flat), iters, conv


def kabsch(src, dst):
    """rigid T (4,4) minimising sum |R s + t - d|^2 (the restatement of pcl::umeyama without scaling), or None when degenerate"""
    src, dst = _c(src), _c(dst)
    T = np.empty(16, np.float32)
    L = lib()
    L.hop_oracle_kabsch.restype = C.c_int
    L.hop_oracle_kabsch.argtypes = [_f32p, _f32p, C.c_int, _f32p]
    return colmajor_to_poses(T)[0] if L.hop_oracle_kabsch(src, dst, len(src), T) else None


def ref_umeyama(src, dst):
    """the reference tree's own Eigen::umeyama (oracle/_ref)"""
    src, dst = _c(src), _c(dst)
    T = np.empty(16, np.float32)
    R = ref()
    R.hop_ref_umeyama.argtypes = [_f32p, _f32p, C.c_int, _f32p]
    R.hop_ref_umeyama.restype = None
    R.hop_ref_umeyama(src, dst, len(src), T)
    return colmajor_to_poses(T)[0]


def refine_by_icp_p2p(s_xyz, m_xyz, poses, max_iter=100, dist=0.01, abs_mse_eps=1e-12, nthreads=0):
    """Utils::runICP(segment, model, T, max_corres_dist) (Utils.cpp:135-164: reciprocal correspondences + SVD) per hypothesis, with
    the pose update of PoseEstimator::refineByICP.  Returns refined (H,4,4), iterations, converged."""
    s_xyz, m_xyz = _c(s_xyz), _c(m_xyz)
    flat = poses_to_colmajor(poses)
    H = len(flat)
    iters, conv = np.zeros(H, np.int32), np.zeros(H, np.int32)
    L = lib()
    L.hop_oracle_refine_by_icp_p2p.restype = None
    L.hop_oracle_refine_by_icp_p2p.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_int,
                                               _i32p, _i32p]
    L.hop_oracle_refine_by_icp_p2p(s_xyz, len(s_xyz), m_xyz, len(m_xyz), flat, H, max_iter, dist, abs_mse_eps, nthreads, iters, conv)
    return colmajor_to_poses(flat), iters, conv


def integral_image_normals(xyz_organized, max_depth_change_factor=0.02, smoothing=10.0):
    """Utils::calNormalIntegralImage(cloud, -1, 0.02, 10, true): xyz_organized (h, w, 3) -> normals (h, w, 3), NaN where PCL writes none"""
    a = np.ascontiguousarray(xyz_organized, np.float32)
    h, w = a.shape[:2]
    out = np.empty((h, w, 3), np.float32)
    L = lib()
    L.hop_oracle_integral_image_normals.restype = None
    L.hop_oracle_integral_image_normals.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_float, _f32p]
    L.hop_oracle_integral_image_normals(a.reshape(-1), w, h, max_depth_change_factor, smoothing, out.reshape(-1))
    return out


def organized_cloud(depth_mm, K):
    """Utils::readDepthImage + convert3dOrganizedRGB (Utils.cpp:36-55, 78-115): (h, w, 3) float32, invalid pixels (0, 0, 0)"""
    d = (depth_mm.astype(np.float32).astype(np.float64) * 0.001).astype(np.float32)
    d = np.where((d.astype(np.float64) > 2.0) | (d.astype(np.float64) < 0.1), np.float32(0), d)
    ok = (d.astype(np.float64) > 0.1) & (d.astype(np.float64) < 2.0)
    h, w = d.shape
    v, u = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    fx, fy, cx, cy = [np.float32(x) for x in K]
    x = ((v - cx) * d) / fx
    y = ((u - cy) * d) / fy
    out = np.stack([x, y, d], -1).astype(np.float32)
    out[~ok] = 0
    return out


def mls(xyz, radius):
    """Utils::calNormalMLS: returns (projected xyz, normals, valid mask) for every input point"""
    a = _c(xyz)
    n = len(a)
    po, no, va = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.zeros(n, np.int32)
    L = lib()
    L.hop_oracle_mls.restype = None
    L.hop_oracle_mls.argtypes = [_f32p, C.c_int, C.c_float, _f32p, _f32p, _i32p]
    L.hop_oracle_mls(a.reshape(-1), n, radius, po.reshape(-1), no.reshape(-1), va)
    return po, no, va.astype(bool)


def select_best(s_xyz, s_nrm, m_xyz, m_nrm, poses, dist=0.001, angle=10.0, weights=None, nthreads=0):
    s_xyz, s_nrm, m_xyz, m_nrm = _c(s_xyz), _c(s_nrm), _c(m_xyz), _c(m_nrm)
    flat = poses_to_colmajor(poses)
    H = len(flat)
    scores = np.zeros(H, np.float32)
    w = None if weights is None else _c(weights).ctypes.data_as(C.c_void_p)
    best = lib().hop_oracle_select_best(s_xyz, s_nrm, len(s_xyz), w, m_xyz, m_nrm, len(m_xyz), flat, H, dist, angle,
                                        nthreads, scores)
    return int(best), scores


def num_threads():
    return int(lib().hop_oracle_num_threads())


# ---- K1 oracle: objFuncPSO over a grid of states (hop_oracle_hand.c) -------------------------------------------------
def hand_overlap(params, f_xyz, f_nrm, nn_xyz, lookup_nrm, w_xyz, thetas, nthreads=0, with_detail=False):
    """params: a ctypes hop_finger_params (hop_b200.FingerParams).  Returns cost (S,) float64 [, detail (S,4)]."""
    f_xyz, f_nrm, nn_xyz, lookup_nrm, w_xyz = _c(f_xyz), _c(f_nrm), _c(nn_xyz), _c(lookup_nrm), _c(w_xyz)
    th = np.ascontiguousarray(thetas, np.float64)
    cost = np.zeros(len(th), np.float64)
    detail = np.zeros((len(th), 4), np.float64) if with_detail else None
    lib().hop_oracle_hand_overlap(C.addressof(params), f_xyz, f_nrm, len(f_xyz), nn_xyz, len(nn_xyz), lookup_nrm, w_xyz, len(w_xyz),
                                  th, len(th), cost, detail.ctypes.data_as(C.c_void_p) if with_detail else None, nthreads)
    return (cost, detail) if with_detail else cost


def finger_property(xyz, num_division, params):
    """FingerProperty restatement: fills params.{num_division,min_z,stride_z,hist_min_y}; returns the bounding box (6,)."""
    bbox = np.zeros(6, np.float32)
    lib().hop_oracle_finger_property(_c(xyz), len(xyz), num_division, C.addressof(params), bbox)
    return bbox


def ref_compute_ppf(p1, n1, p2, n2):
    key = np.zeros(4, np.int32)
    ref().hop_ref_compute_ppf(_c(p1), _c(n1), _c(p2), _c(n2), key)
    return key


def ref_cluster_poses(poses, scores, angle_diff, dist_diff, symmetry_deg=(360.0, 360.0, 360.0), ids=None):
    """clusterPoses restated on the reference tree's own Eigen (oracle/_ref/ref_cluster.cpp): kept indices in cluster order."""
    flat = poses_to_colmajor(poses)
    sc = _c(scores)
    idv = None if ids is None else np.ascontiguousarray(ids, np.int32)
    keep = np.zeros(max(len(flat), 1), np.int32)
    n = ref().hop_ref_cluster_poses(flat, sc, None if idv is None else idv.ctypes.data_as(C.c_void_p), len(flat), angle_diff, dist_diff,
                                    np.ascontiguousarray(symmetry_deg, np.float64), keep)
    return keep[:n].copy()


# ---- physics pruning ---------------------------------------------------------------------------
class CollisionParams(C.Structure):
    """same layout as hop_collision_params (include/hop_c_api.h)"""
    _fields_ = [("cam2handbase", C.c_float * 16), ("model_center", C.c_float * 3), ("ob_diameter", C.c_float),
                ("collision_dist", C.c_float), ("inside_ob_dist", C.c_float), ("non_touch_dist", C.c_float),
                ("collision_finger_dist", C.c_float), ("collision_finger_volume_ratio", C.c_float), ("finger_status", C.c_int32 * 4)]


def collision_params(d, cls=CollisionParams):
    p = cls()
    p.cam2handbase[:] = np.asarray(d["cam2handbase"], np.float32).T.reshape(-1).tolist()   # column-major
    p.model_center[:] = [float(v) for v in d["model_center"]]
    for k in ("ob_diameter", "collision_dist", "inside_ob_dist", "non_touch_dist", "collision_finger_dist", "collision_finger_volume_ratio"):
        setattr(p, k, float(d[k]))
    p.finger_status[:] = [int(v) for v in d["finger_status"]]
    return p


def signed_distance(pts, V, F):
    """restated igl::signed_distance (pseudonormal): returns S, closest face, closest point"""
    L = lib()
    pts, V, F = _c(pts), _c(V), np.ascontiguousarray(F, np.int32)
    S, I, Cp = np.empty(len(pts), np.float32), np.empty(len(pts), np.int32), np.empty((len(pts), 3), np.float32)
    L.hop_oracle_signed_distance.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, _f32p, _i32p, _f32p]
    rc = L.hop_oracle_signed_distance(pts, len(pts), V, len(V), F, len(F), S, I, Cp)
    assert rc == 0
    return S, I, Cp


def ref_signed_distance(pts, V, F):
    """the reference's vendored libigl (compiled where it lies): S, I, C, N of igl::signed_distance"""
    R = ref()
    pts, V, F = _c(pts), _c(V), np.ascontiguousarray(F, np.int32)
    n = len(pts)
    S, I, Cp, N = np.empty(n, np.float32), np.empty(n, np.int32), np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
    R.hop_ref_signed_distance.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, _f32p, _i32p, _f32p, _f32p]
    R.hop_ref_signed_distance(pts, n, V, len(V), F, len(F), S, I, Cp, N)
    return S, I, Cp, N


def reject_by_collision(case, nthreads=0, with_ambiguous=False):
    """restated PoseEstimator::rejectByCollisionOrNonTouching over a synth.make_collision_case dict: keep, reason, diag"""
    L = lib()
    H = len(case["poses"])
    fV = _c(np.concatenate(case["finger_V"])) if sum(len(v) for v in case["finger_V"]) else np.zeros((1, 3), np.float32)
    fF = np.ascontiguousarray(np.concatenate(case["finger_F"]), np.int32) if sum(len(f) for f in case["finger_F"]) else np.zeros((1, 3), np.int32)
    fP = _c(np.concatenate(case["finger_pts"])) if sum(len(v) for v in case["finger_pts"]) else np.zeros((1, 3), np.float32)
    fnv = np.array([len(v) for v in case["finger_V"]], np.int32)
    fnf = np.array([len(f) for f in case["finger_F"]], np.int32)
    fn = np.array([len(v) for v in case["finger_pts"]], np.int32)
    keep, reason, diag = np.empty(H, np.int32), np.empty(H, np.int32), np.empty((H, 10), np.float32)
    p = collision_params(case["params"])
    amb = np.zeros(H, np.int32)
    L.hop_oracle_reject_by_collision_amb.argtypes = [_f32p, C.c_int, _i32p, C.c_int, _f32p, _i32p, _i32p, _i32p, _f32p, _i32p, _f32p, C.c_int,
                                                     _f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.POINTER(CollisionParams), _i32p, _i32p, _f32p,
                                                     C.c_void_p]
    rc = L.hop_oracle_reject_by_collision_amb(_c(case["obj_V"]), len(case["obj_V"]), np.ascontiguousarray(case["obj_F"], np.int32), len(case["obj_F"]),
                                          fV, fnv, fF, fnf, fP, fn, _c(case["scene_xyz"]), len(case["scene_xyz"]), _c(case["hand_xyz"]),
                                          len(case["hand_xyz"]), _c(case["model_xyz"]), len(case["model_xyz"]),
                                          poses_to_colmajor(case["poses"]), H, C.byref(p), keep, reason, diag,
                                          amb.ctypes.data if with_ambiguous else None)
    assert rc == 0
    return (keep, reason, diag, amb) if with_ambiguous else (keep, reason, diag)


def signed_distance_ambiguous(pts, V, F):
    """S and per-point flags: 1 where the sign of the distance is a coin toss between tied faces / candidate normals
    (hop_oracle_sdf.c: sdf_point_amb) -- what another summation order, an FMA or igl's AABB traversal may resolve the other way"""
    L = lib()
    pts, V, F = _c(pts), _c(V), np.ascontiguousarray(F, np.int32)
    S, A = np.empty(len(pts), np.float32), np.empty(len(pts), np.int32)
    L.hop_oracle_signed_distance_amb.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, _f32p, _i32p]
    assert L.hop_oracle_signed_distance_amb(pts, len(pts), V, len(V), F, len(F), S, A) == 0
    return S, A


# ---- render-based rejection --------------------------------------------------------------------
class RenderParams(C.Structure):
    """same layout as hop_render_params (include/hop_c_api.h)"""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("z_near", C.c_float), ("z_far", C.c_float), ("roi_weight", C.c_float), ("keep_ratio", C.c_float)]


def render_params(**kw):
    p = RenderParams(616.596, 616.596, 307.628, 239.687, 640, 480, 0.1, 2.0, 2.0, 0.3)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def _mesh_args(V, F):
    if V is None or len(V) == 0 or F is None or len(F) == 0:
        return np.zeros((1, 3), np.float32), 0, np.zeros((1, 3), np.int32), 0
    V, F = _c(V), np.ascontiguousarray(F, np.int32)
    return V, len(V), F, len(F)


def render_depth(p, hand_V, hand_F, obj_V, obj_F, pose):
    """restated Renderer::doRender: (depth (h,w) metres, mask (h,w) uint8 = object pixels)"""
    L = lib()
    hV, hnv, hF, hnf = _mesh_args(hand_V, hand_F)
    oV, onv, oF, onf = _mesh_args(obj_V, obj_F)
    depth, mask = np.empty((p.height, p.width), np.float32), np.empty((p.height, p.width), np.uint8)
    L.hop_oracle_render_depth.argtypes = [C.POINTER(RenderParams), _f32p, C.c_int, _i32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, _f32p, _f32p, _u8p]
    rc = L.hop_oracle_render_depth(C.byref(p), hV, hnv, hF, hnf, oV, onv, oF, onf, poses_to_colmajor(np.asarray(pose).reshape(1, 4, 4)), depth, mask)
    assert rc == 0
    return depth, mask


def reject_by_render(p, depth_m, hand_V, hand_F, obj_V, obj_F, poses):
    """restated PoseEstimator::rejectByRender: (wrong_ratio (H,), kept indices in the reference's output order)"""
    L = lib()
    hV, hnv, hF, hnf = _mesh_args(hand_V, hand_F)
    oV, onv, oF, onf = _mesh_args(obj_V, obj_F)
    flat = poses_to_colmajor(poses)
    H = len(flat)
    wr, order, nk = np.empty(H, np.float32), np.zeros(max(H, 1), np.int32), C.c_int32(0)
    L.hop_oracle_reject_by_render.argtypes = [C.POINTER(RenderParams), _f32p, _f32p, C.c_int, _i32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, _f32p,
                                              C.c_int, _f32p, _i32p, C.POINTER(C.c_int32)]
    rc = L.hop_oracle_reject_by_render(C.byref(p), _c(depth_m), hV, hnv, hF, hnf, oV, onv, oF, onf, flat, H, wr, order, C.byref(nk))
    assert rc == 0
    return wr, order[: nk.value].copy()


# ---- hand-point removal (Hand.cpp:781-888) ------------------------------------------------------------------------
class HandRemovalParams(C.Structure):
    """same layout as hop_hand_removal_params (include/hop_c_api.h)"""
    _fields_ = [("cam_in_handbase", C.c_float * 16), ("handbase_in_cam", C.c_float * 16), ("handbase_in_finger_1_2", C.c_float * 16),
                ("handbase_in_finger_2_2", C.c_float * 16), ("min_z", C.c_float), ("dist_thres_sq", C.c_float)]


def remove_hand_points(xyz, nrm, links, link_kind, params):
    """restated HandT42::removeSurroundingPointsAndAssignProbability; params: any structure with hop_hand_removal_params' layout"""
    L = lib()
    xyz, nrm = _c(xyz), _c(nrm)
    n = len(xyz)
    lk = _c(np.concatenate([np.asarray(l, np.float32).reshape(-1, 3) for l in links])) if len(links) and sum(len(l) for l in links) else np.zeros((1, 3), np.float32)
    ln = np.array([len(l) for l in links], np.int32) if len(links) else np.zeros(1, np.int32)
    kinds = np.ascontiguousarray(link_kind, np.int32) if len(links) else np.zeros(1, np.int32)
    ox, on, oc = np.empty((max(n, 1), 3), np.float32), np.empty((max(n, 1), 3), np.float32), np.empty(max(n, 1), np.float32)
    L.hop_oracle_remove_hand_points.argtypes = [_f32p, _f32p, C.c_int, _f32p, _i32p, _i32p, C.c_int, C.c_void_p, _f32p, _f32p, _f32p]
    m = L.hop_oracle_remove_hand_points(xyz if n else np.zeros((1, 3), np.float32), nrm if n else np.zeros((1, 3), np.float32), n, lk, ln, kinds, len(links),
                                        C.cast(C.byref(params), C.c_void_p), ox, on, oc)
    assert m >= 0
    return ox[:m].copy(), on[:m].copy(), oc[:m].copy()


def adjust_hand_height(hand_xyz, hand_nrm, scene_xyz, scene_nrm, heights):
    """restated HandT42::adjustHandHeight: (counts per height, chosen index or -1); scene in the hand-base frame"""
    L = lib()
    hs = np.ascontiguousarray(heights, np.float32)
    counts = np.zeros(len(hs), np.int32)
    L.hop_oracle_adjust_hand_height.argtypes = [_f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int, _i32p]
    best = L.hop_oracle_adjust_hand_height(_c(hand_xyz), _c(hand_nrm), len(hand_xyz), _c(scene_xyz), _c(scene_nrm), len(scene_xyz), hs, len(hs), counts)
    return counts, int(best)


# ---- the cloud filters of Hand::setCurScene (Hand.cpp:279-334), numpy restatements (float32 distances in PCL/FLANN's order) ----
def _d2_matrix(xyz):
    p = np.asarray(xyz, np.float32)
    dx = p[None, :, 0] - p[:, None, 0]
    dy = p[None, :, 1] - p[:, None, 1]
    dz = p[None, :, 2] - p[:, None, 2]
    return (dx * dx + dy * dy) + dz * dz           # float32, unfused, left to right


def radius_outlier_removal(xyz, radius, min_neighbors):
    """pcl::RadiusOutlierRemoval (dense input): keep mask; a point stays with more than min_neighbors points (itself included) in d^2 <= float(r^2)"""
    r2 = np.float32(float(radius) * float(radius))
    return (_d2_matrix(xyz) <= r2).sum(1) >= min_neighbors + 1


def statistical_outlier_removal(xyz, mean_k, stddev_mul):
    """pcl::StatisticalOutlierRemoval: (keep mask, mean distances); double sums in ascending / point order like the reference"""
    d2 = np.sort(_d2_matrix(xyz), axis=1)[:, 1:mean_k + 1]
    s = np.zeros(len(d2), np.float64)
    for k in range(d2.shape[1]):
        s = s + np.sqrt(d2[:, k]).astype(np.float64)
    dist = (s / float(mean_k)).astype(np.float32)
    n = len(dist)
    total = np.cumsum(dist.astype(np.float64))[-1]
    sq = np.cumsum(dist.astype(np.float64) * dist.astype(np.float64))[-1]
    mean = total / n
    var = (sq - total * total / n) / (n - 1)
    thr = mean + float(np.float32(stddev_mul)) * np.sqrt(var)
    return dist.astype(np.float64) <= thr, dist
