"""ctypes binding of include/hop_c_api.h (the drop-in boundary).  No torch types, plain pointers and sizes."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
_HEADER = os.path.join(_ROOT, "include", "hop_c_api.h")


class HopError(RuntimeError):
    pass


class IcpParams(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("angle_deg", C.c_float), ("max_dist", C.c_float), ("abs_mse_eps", C.c_double),
                ("mode", C.c_int32), ("solver", C.c_int32), ("team_warps", C.c_int32), ("pipeline", C.c_int32)]


class FrameParams(C.Structure):
    """hop_frame_params (include/hop_c_api.h): the per-frame front end of main_realdata_auto.cpp:54-96,144-181"""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("leaf_dense", C.c_float),
                ("cam_in_handbase", C.c_float * 16), ("handbase_in_cam", C.c_float * 16), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
                ("normal_radius", C.c_float), ("leaf_object", C.c_float), ("viewpoint", C.c_float * 3)]


class LcpParams(C.Structure):
    _fields_ = [("dist", C.c_float), ("angle_deg", C.c_float), ("use_normal", C.c_int32), ("use_dot_score", C.c_int32),
                ("use_reciprocal", C.c_int32), ("team_warps", C.c_int32)]


HOP_MAX_FINGER_BINS = 32


class FingerParams(C.Structure):
    """hop_finger_params: what objFuncPSO reads from optim::ArgPasser and the YAML (Hand.cpp:10-178)."""
    _fields_ = [("model2handbase", C.c_float * 16), ("finger_out2parent", C.c_float * 16), ("tip1_local", C.c_float * 3),
                ("tip2_local", C.c_float * 3), ("pair_tip1_y", C.c_float), ("pair_tip2_y", C.c_float), ("palm_side", C.c_int32),
                ("right_side", C.c_int32), ("gripper_min_dist", C.c_float), ("dist_thres", C.c_float), ("normal_angle_deg", C.c_float),
                ("check_normal", C.c_int32), ("num_division", C.c_int32), ("min_z", C.c_float), ("stride_z", C.c_float),
                ("hist_min_y", C.c_float * HOP_MAX_FINGER_BINS), ("max_outter_pts", C.c_int32), ("outter_pt_dist", C.c_float),
                ("outter_pt_dist_weight", C.c_float)]


class S4pcsOptions(C.Structure):
    """hop_s4pcs_options: what PoseEstimator::runSuper4pcs sets (PoseEstimator.cpp:66-73)."""
    _fields_ = [("sample_size", C.c_int32), ("overlap", C.c_float), ("delta", C.c_float), ("dispersion", C.c_float),
                ("success_quadrilaterals", C.c_int32), ("max_normal_difference", C.c_float), ("max_color_distance", C.c_float),
                ("max_trials", C.c_int32), ("random_seed", C.c_uint32), ("keep_intermediates", C.c_int32)]


class CollisionParams(C.Structure):
    """hop_collision_params (include/hop_c_api.h)"""
    _fields_ = [("cam2handbase", C.c_float * 16), ("model_center", C.c_float * 3), ("ob_diameter", C.c_float),
                ("collision_dist", C.c_float), ("inside_ob_dist", C.c_float), ("non_touch_dist", C.c_float),
                ("collision_finger_dist", C.c_float), ("collision_finger_volume_ratio", C.c_float), ("finger_status", C.c_int32 * 4)]


class HandRemovalParams(C.Structure):
    """hop_hand_removal_params (include/hop_c_api.h)"""
    _fields_ = [("cam_in_handbase", C.c_float * 16), ("handbase_in_cam", C.c_float * 16), ("handbase_in_finger_1_2", C.c_float * 16),
                ("handbase_in_finger_2_2", C.c_float * 16), ("min_z", C.c_float), ("dist_thres_sq", C.c_float)]


class RenderParams(C.Structure):
    """hop_render_params (include/hop_c_api.h)"""
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("z_near", C.c_float), ("z_far", C.c_float), ("roi_weight", C.c_float), ("keep_ratio", C.c_float)]


class PoseRec(C.Structure):
    _fields_ = [("pose", C.c_float * 16), ("score", C.c_float), ("id", C.c_int32), ("frame", C.c_int32), ("pad", C.c_int32)]


POSE_REC_DTYPE = np.dtype([("pose", np.float32, (16,)), ("score", np.float32), ("id", np.int32), ("frame", np.int32),
                           ("pad", np.int32)])
assert POSE_REC_DTYPE.itemsize == 80 and C.sizeof(PoseRec) == 80


def lib_path():
    return os.path.join(_CSRC, "libhop.so")


def build_library(verbose=False):
    """Compile libhop.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _CSRC, "-j4"], capture_output=True, text=True)
    if r.returncode != 0:
        raise HopError("libhop build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return lib_path()


def declared_symbols():
    """Every function the public header declares (used by the CPU tests to check the exports)."""
    txt = open(_HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hop_[a-z0-9_]+)\s*\(", txt)))


_lib = None
_vp = C.c_void_p


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise HopError(f"{path} is missing: run __graft_entry__.build() (make -C {_CSRC}); there is no CPU fallback")
    L = C.CDLL(path)
    L.hop_create.argtypes = [C.c_int, C.POINTER(_vp)]
    L.hop_destroy.argtypes = [_vp]
    L.hop_destroy.restype = None
    L.hop_last_error.argtypes = [_vp]
    L.hop_last_error.restype = C.c_char_p
    L.hop_set_stream.argtypes = [_vp, _vp]
    L.hop_sync.argtypes = [_vp]
    L.hop_launch_count.argtypes = [_vp]
    L.hop_launch_count.restype = C.c_int64
    L.hop_profile_enable.argtypes = [_vp, C.c_int]
    L.hop_profile_read.argtypes = [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.hop_default_icp_params.argtypes = [C.POINTER(IcpParams)]
    L.hop_default_icp_params.restype = None
    L.hop_default_lcp_params.argtypes = [C.POINTER(LcpParams)]
    L.hop_default_lcp_params.restype = None
    L.hop_malloc.argtypes = [_vp, C.c_size_t, C.POINTER(_vp)]
    L.hop_free.argtypes = [_vp, _vp]
    L.hop_host_alloc.argtypes = [_vp, C.c_size_t, C.POINTER(_vp)]
    L.hop_host_free.argtypes = [_vp, _vp]
    L.hop_memcpy_h2d.argtypes = [_vp, _vp, _vp, C.c_size_t]
    L.hop_memcpy_d2h.argtypes = [_vp, _vp, _vp, C.c_size_t]
    L.hop_cloud_upload.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(_vp)]
    L.hop_cloud_update.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int]
    L.hop_cloud_free.argtypes = [_vp, _vp]
    L.hop_cloud_size.argtypes = [_vp]
    L.hop_cloud_prepare_nn.argtypes = [_vp, _vp, C.c_float, C.c_float, C.POINTER(C.c_int64)]
    L.hop_cloud_prepare_nn_async.argtypes = [_vp, _vp, C.c_float, C.c_float]
    L.hop_lcp_prepare_scene_async.argtypes = [_vp, _vp, _vp, C.c_int]
    L.hop_cloud_drop_nn.argtypes = [_vp, _vp]
    L.hop_cloud_hint_static.argtypes = [_vp, _vp, C.c_int]
    L.hop_cloud_nn_query.argtypes = [_vp, _vp, C.c_float, _vp, C.c_int, _vp, _vp]
    L.hop_icp_refine.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(IcpParams), _vp, _vp]
    L.hop_icp_refine_dev.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(IcpParams), _vp, _vp]
    L.hop_lcp_score.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(LcpParams), C.c_int, _vp]
    L.hop_lcp_score_dev.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(LcpParams), C.c_int, _vp]
    L.hop_verify_lcp.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp]
    L.hop_verify_lcp_dev.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_float, _vp, _vp, _vp, _vp]
    L.hop_default_s4pcs_options.argtypes = [C.POINTER(S4pcsOptions)]
    L.hop_default_s4pcs_options.restype = None
    L.hop_s4pcs_plan_create.argtypes = [_vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(S4pcsOptions), C.POINTER(_vp)]
    L.hop_s4pcs_plan_destroy.argtypes = [_vp]
    L.hop_s4pcs_plan_destroy.restype = None
    L.hop_s4pcs_plan_sizes.argtypes = [_vp, _vp]
    L.hop_s4pcs_plan_get.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.hop_s4pcs_plan_intermediates.argtypes = [_vp, _vp, _vp, _vp]
    L.hop_compute_ppf.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.hop_compute_ppf.restype = None
    L.hop_super4pcs_run.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _vp]
    L.hop_cluster_poses.argtypes = [_vp, _vp, _vp, C.c_int, C.c_float, C.c_float, _vp, _vp, _vp]
    L.hop_default_frame_params.argtypes = [C.POINTER(FrameParams)]
    L.hop_default_frame_params.restype = None
    L.hop_frame_to_scene.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(FrameParams), C.POINTER(_vp), _vp]
    L.hop_cloud_download.argtypes = [_vp, _vp, _vp, _vp, _vp]
    L.hop_cluster_poses_gpu.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_float, _vp, _vp, _vp]
    L.hop_hand_overlap.argtypes = [_vp, _vp, _vp, _vp, _vp, C.POINTER(FingerParams), _vp, C.c_int, _vp, _vp]
    L.hop_hand_overlap_dev.argtypes = [_vp, _vp, _vp, _vp, _vp, C.POINTER(FingerParams), _vp, _vp, C.c_int, _vp, _vp]
    L.hop_select_topk_dev.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int32, C.c_int32, _vp]
    L.hop_select_topk.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int32, C.c_int32, _vp]
    L.hop_mesh_upload.argtypes = [_vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(_vp)]
    L.hop_mesh_free.argtypes = [_vp, _vp]
    L.hop_sdf_query.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]
    L.hop_reject_by_collision.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(CollisionParams), _vp, _vp, _vp]
    L.hop_reject_by_collision_dev.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(CollisionParams), _vp, _vp, _vp]
    L.hop_cloud_voxel_grid.argtypes = [_vp, _vp, C.c_float, C.POINTER(_vp)]
    L.hop_cloud_transform.argtypes = [_vp, _vp, _vp, C.POINTER(_vp)]
    L.hop_cloud_pass_through.argtypes = [_vp, _vp, C.c_int, C.c_float, C.c_float, C.POINTER(_vp)]
    L.hop_cloud_handbase_region.argtypes = [_vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(_vp)]
    L.hop_cloud_radius_outlier_removal.argtypes = [_vp, _vp, C.c_float, C.c_int, C.POINTER(_vp)]
    L.hop_cloud_statistical_outlier_removal.argtypes = [_vp, _vp, C.c_int, C.c_float, C.POINTER(_vp)]
    L.hop_adjust_hand_height.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp]
    L.hop_remove_hand_points.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.POINTER(HandRemovalParams), C.POINTER(_vp)]
    L.hop_default_render_params.argtypes = [C.POINTER(RenderParams)]
    L.hop_default_render_params.restype = None
    L.hop_render_scene_create.argtypes = [_vp, C.POINTER(RenderParams), _vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(_vp)]
    L.hop_render_scene_destroy.argtypes = [_vp, _vp]
    L.hop_render_depth.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]
    L.hop_reject_by_render.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]
    L.hop_refine_score_select.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(IcpParams), C.POINTER(LcpParams), C.c_int, C.c_int,
                                          _vp, _vp, _vp, _vp, _vp]
    L.hop_refine_score_select_dev.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(IcpParams), C.POINTER(LcpParams), C.c_int, C.c_int,
                                              C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]
    L.hop_s4pcs_plan_create_gpu.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(S4pcsOptions), C.POINTER(_vp)]
    L.hop_ppf_table_build.argtypes = [_vp, _vp, _vp, C.c_int, _vp, C.c_int, C.POINTER(C.c_int32)]
    L.hop_debug_lm_solve.argtypes = [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp]
    L.hop_frame_organized.argtypes = [_vp, _vp, C.c_int, C.c_int, C.POINTER(FrameParams), C.c_float, C.c_float, C.POINTER(_vp)]
    L.hop_cloud_mls.argtypes = [_vp, _vp, C.c_float, C.POINTER(_vp)]
    L.hop_comm_unique_id.argtypes = [_vp]
    L.hop_comm_init.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.hop_comm_destroy.argtypes = [_vp]
    L.hop_comm_rank.argtypes = [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.hop_gather_winners.argtypes = [_vp, _vp, C.c_int, _vp]
    L.hop_gather_winners_dev.argtypes = [_vp, _vp, C.c_int, _vp, C.c_int]
    L.hop_gather_wait.argtypes = [_vp, C.c_int]
    for name in declared_symbols():
        fn = getattr(L, name)  # raises AttributeError when an export is missing
        if fn.restype is C.c_int and name not in ("hop_cloud_size",):
            pass
    _lib = L
    return L


def _f32(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape_last is not None and (a.ndim != 2 or a.shape[1] != shape_last):
        raise ValueError(f"expected (N,{shape_last}) array, got {a.shape}")
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def comm_unique_id():
    """rank 0: the ncclUniqueId (128 bytes) every rank passes to Context.comm_init"""
    L = load_library()
    buf = (C.c_char * 128)()
    rc = L.hop_comm_unique_id(buf)
    if rc != 0:
        raise HopError(f"hop_comm_unique_id failed ({rc}): is libnccl.so.2 on the loader path?")
    return bytes(buf)


def poses_to_colmajor(poses):
    """(H,4,4) numpy (row-major, math convention) -> (H,16) column-major float32 = Eigen::Matrix4f::data()."""
    p = np.asarray(poses, dtype=np.float32).reshape(-1, 4, 4)
    return np.ascontiguousarray(p.transpose(0, 2, 1).reshape(-1, 16))


def colmajor_to_poses(flat):
    return np.asarray(flat, dtype=np.float32).reshape(-1, 4, 4).transpose(0, 2, 1).copy()


def s4pcs_options(**kw):
    o = S4pcsOptions()
    load_library().hop_default_s4pcs_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def compute_ppf(p1, n1, p2, n2):
    """gr::computePPF of one point pair (host function of libhop): 4-int key."""
    key = np.zeros(4, np.int32)
    a = [np.ascontiguousarray(v, np.float32) for v in (p1, n1, p2, n2)]
    load_library().hop_compute_ppf(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(key))
    return key


def cluster_poses(poses, scores, angle_diff_deg, dist_diff, symmetry_deg=(360.0, 360.0, 360.0), ids=None):
    """PoseEstimator::clusterPoses on the host (libhop, no GPU): returns the indices of the kept hypotheses in cluster order."""
    flat = poses_to_colmajor(poses)
    sc = _f32(scores)
    n = len(flat)
    idv = None if ids is None else np.ascontiguousarray(ids, np.int32)
    sym = np.ascontiguousarray(symmetry_deg, np.float32)
    keep = np.zeros(max(n, 1), np.int32)
    nk = C.c_int32(0)
    rc = load_library().hop_cluster_poses(_ptr(flat), _ptr(sc), _ptr(idv), n, angle_diff_deg, dist_diff, _ptr(sym), _ptr(keep), C.byref(nk))
    if rc != 0:
        raise HopError(f"hop_cluster_poses failed ({rc})")
    return keep[: nk.value].copy()


class S4pcsPlan:
    """hop_s4pcs_plan*: the plan of one Super4PCS registration (sampling + all bases).  ctx = None: entirely on the host (no GPU
    needed); with a Context the planner's PPF-membership scans are answered by the device (hop_s4pcs_plan_create_gpu), same plan."""

    def __init__(self, P_xyz, P_nrm, P_prob, Q_xyz, Q_nrm, ppf_keys, options=None, ctx=None):
        self.L = load_library()
        self.options = options or s4pcs_options()
        P_xyz, P_nrm, Q_xyz, Q_nrm = _f32(P_xyz, 3), _f32(P_nrm, 3), _f32(Q_xyz, 3), _f32(Q_nrm, 3)
        prob = None if P_prob is None else _f32(P_prob)
        keys = np.ascontiguousarray(ppf_keys, np.int32).reshape(-1, 4)
        h = _vp()
        if ctx is None:
            rc = self.L.hop_s4pcs_plan_create(_ptr(P_xyz), _ptr(P_nrm), _ptr(prob), len(P_xyz), _ptr(Q_xyz), _ptr(Q_nrm), len(Q_xyz), _ptr(keys),
                                              len(keys), C.byref(self.options), C.byref(h))
        else:
            rc = self.L.hop_s4pcs_plan_create_gpu(ctx.h, _ptr(P_xyz), _ptr(P_nrm), _ptr(prob), len(P_xyz), _ptr(Q_xyz), _ptr(Q_nrm), len(Q_xyz),
                                                  _ptr(keys), len(keys), C.byref(self.options), C.byref(h))
        if rc != 0:
            raise HopError(f"hop_s4pcs_plan_create failed ({rc})")
        self.h = h

    def sizes(self):
        s = np.zeros(6, np.int32)
        self.L.hop_s4pcs_plan_sizes(self.h, _ptr(s))
        return dict(nP=int(s[0]), nQ=int(s[1]), trials=int(s[2]), pairs=int(s[3]), quads=int(s[4]), trials_executed=int(s[5]))

    def get(self):
        z = self.sizes()
        Pc, Qc = np.zeros((z["nP"], 3), np.float32), np.zeros((z["nQ"], 3), np.float32)
        q_ids, cen, misc = np.zeros(z["nQ"], np.int32), np.zeros(6, np.float32), np.zeros(2, np.float32)
        ti, tf = np.zeros((z["trials"], 5), np.int32), np.zeros((z["trials"], 4), np.float32)
        self.L.hop_s4pcs_plan_get(self.h, _ptr(Pc), _ptr(Qc), _ptr(q_ids), _ptr(cen), _ptr(misc), _ptr(ti), _ptr(tf))
        return dict(Pc=Pc, Qc=Qc, q_ids=q_ids, centroid_P=cen[:3].copy(), centroid_Q=cen[3:].copy(), diameter=float(misc[0]),
                    ratio=float(misc[1]), base_ok=ti[:, 0].copy(), bases=ti[:, 1:].copy(), inv=tf[:, :2].copy(), dist=tf[:, 2:].copy())

    def intermediates(self):
        z = self.sizes()
        tr, pairs, quads = np.zeros((z["trials"], 6), np.int32), np.zeros((z["pairs"], 2), np.int32), np.zeros((z["quads"], 4), np.int32)
        self.L.hop_s4pcs_plan_intermediates(self.h, _ptr(tr), _ptr(pairs), _ptr(quads))
        return tr, pairs, quads

    def close(self):
        if getattr(self, "h", None):
            self.L.hop_s4pcs_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Cloud:
    """Device-resident point cloud handle (hop_cloud*)."""

    def __init__(self, ctx, handle, n):
        self.ctx, self.handle, self.n = ctx, handle, n

    def update(self, xyz, nrm=None, prob=None):
        xyz = _f32(xyz, 3)
        nrm = None if nrm is None else _f32(nrm, 3)
        prob = None if prob is None else _f32(prob)
        self.ctx._check(self.ctx.L.hop_cloud_update(self.ctx.h, self.handle, _ptr(xyz), _ptr(nrm), _ptr(prob), len(xyz)))
        self.n = len(xyz)

    def download(self):
        """(xyz (n,3), normals (n,3), confidence (n,)) of the device cloud"""
        n = int(self.ctx.L.hop_cloud_size(self.handle))
        xyz, nrm, prob = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros(n, np.float32)
        self.ctx._check(self.ctx.L.hop_cloud_download(self.ctx.h, self.handle, _ptr(xyz), _ptr(nrm), _ptr(prob)))
        return xyz, nrm, prob

    def prepare_nn(self, radius, voxel=0.0):
        stats = (C.c_int64 * 4)()
        self.ctx._check(self.ctx.L.hop_cloud_prepare_nn(self.ctx.h, self.handle, radius, voxel, stats))
        return {"voxels": stats[0], "entries": stats[1], "max_list": stats[2], "bytes": stats[3]}

    def prepare_nn_async(self, radius, voxel=0.0):
        """hop_cloud_prepare_nn_async: the grid is built on the context's second stream while the main stream goes on."""
        self.ctx._check(self.ctx.L.hop_cloud_prepare_nn_async(self.ctx.h, self.handle, radius, voxel))

    def prepare_lcp_scene(self, lcp_params, batch=0):
        """hop_lcp_prepare_scene_async: the scene grid hop_lcp_score's reciprocal term needs, built ahead on the second stream; `batch` =
        how many hypotheses will be scored against this frame (chooses the voxel edge)."""
        self.ctx._check(self.ctx.L.hop_lcp_prepare_scene_async(self.ctx.h, self.handle, C.byref(lcp_params), int(batch)))

    def hint_static(self, is_static=True):
        """hop_cloud_hint_static: the cloud is a model (contents stay); grids built afterwards may be finer.  Returns self."""
        self.ctx._check(self.ctx.L.hop_cloud_hint_static(self.ctx.h, self.handle, int(is_static)))
        return self

    def drop_nn(self):
        self.ctx._check(self.ctx.L.hop_cloud_drop_nn(self.ctx.h, self.handle))

    def nn_query(self, radius, queries):
        q = _f32(queries, 3)
        idx = np.empty(len(q), np.int32)
        d2 = np.empty(len(q), np.float32)
        self.ctx._check(self.ctx.L.hop_cloud_nn_query(self.ctx.h, self.handle, radius, _ptr(q), len(q), _ptr(idx), _ptr(d2)))
        return idx, d2

    def free(self):
        if self.handle:
            self.ctx.L.hop_cloud_free(self.ctx.h, self.handle)
            self.handle = None

    # -- the filters of Hand::setCurScene (Hand.cpp:279-334): each returns a new device cloud, input order kept
    def _new(self, fn, *args):
        h = _vp()
        self.ctx._check(fn(self.ctx.h, self.handle, *args, C.byref(h)))
        return Cloud(self.ctx, h, int(self.ctx.L.hop_cloud_size(h)))

    def voxel_grid(self, leaf):
        return self._new(self.ctx.L.hop_cloud_voxel_grid, C.c_float(leaf))

    def transform(self, T):
        flat = poses_to_colmajor(np.asarray(T, np.float32).reshape(1, 4, 4))
        return self._new(self.ctx.L.hop_cloud_transform, _ptr(flat))

    def pass_through(self, axis, lo, hi):
        return self._new(self.ctx.L.hop_cloud_pass_through, {"x": 0, "y": 1, "z": 2}.get(axis, axis), C.c_float(lo), C.c_float(hi))

    def handbase_region(self, y1, z1, y2, z2):
        return self._new(self.ctx.L.hop_cloud_handbase_region, C.c_float(y1), C.c_float(z1), C.c_float(y2), C.c_float(z2))

    def radius_outlier_removal(self, radius, min_neighbors):
        return self._new(self.ctx.L.hop_cloud_radius_outlier_removal, C.c_float(radius), int(min_neighbors))

    def mls(self, radius):
        """Utils::calNormalMLS: the points with >= 3 neighbours projected onto their MLS surfaces, with the surface normals"""
        return self._new(self.ctx.L.hop_cloud_mls, C.c_float(radius))

    def statistical_outlier_removal(self, mean_k, stddev_mul):
        return self._new(self.ctx.L.hop_cloud_statistical_outlier_removal, int(mean_k), C.c_float(stddev_mul))


class Mesh:
    """Device-resident triangle mesh with igl's pseudonormals (hop_mesh*)."""

    def __init__(self, ctx, handle, nv, nf):
        self.ctx, self.handle, self.nv, self.nf = ctx, handle, nv, nf

    def free(self):
        if self.handle:
            self.ctx.L.hop_mesh_free(self.ctx.h, self.handle)
            self.handle = None


class RenderScene:
    """hop_render_scene*: the real depth image + the hand meshes of one frame, rasterised once"""

    def __init__(self, ctx, handle, params):
        self.ctx, self.handle, self.params = ctx, handle, params

    def free(self):
        if self.handle:
            self.ctx.L.hop_render_scene_destroy(self.ctx.h, self.handle)
            self.handle = None


class Context:
    """hop_ctx*: one per process/GPU.  Raises HopError when no B200-class device is present (no CPU fallback)."""

    def __init__(self, device=0):
        self.L = load_library()
        h = _vp()
        rc = self.L.hop_create(device, C.byref(h))
        if rc != 0:
            raise HopError(f"hop_create({device}) failed ({rc}): {self.L.hop_last_error(None).decode()}")
        self.h = h
        self.device = device

    def _check(self, rc):
        if rc != 0:
            raise HopError(f"libhop error {rc}: {self.L.hop_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.L.hop_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def set_stream(self, cuda_stream_handle):
        self._check(self.L.hop_set_stream(self.h, _vp(cuda_stream_handle) if cuda_stream_handle else None))

    def sync(self):
        self._check(self.L.hop_sync(self.h))

    def launch_count(self):
        return int(self.L.hop_launch_count(self.h))

    PROF_KINDS = {"icp_correspond": 0, "icp_solve": 1, "lcp_score": 2, "nn_build": 3, "topk": 4, "verify_lcp": 5, "hand_overlap": 6, "icp_fused": 7, "s4pcs_pairs": 8, "s4pcs_join": 9, "cluster": 10, "frame": 11, "sdf": 12, "render": 13}

    def profile_enable(self, on=True):
        self._check(self.L.hop_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{kernel family: (total ms, spans)} measured with CUDA events on the context's stream (synchronises)."""
        out = {}
        for name, k in self.PROF_KINDS.items():
            ms, n = C.c_double(), C.c_int64()
            self._check(self.L.hop_profile_read(self.h, k, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def malloc(self, nbytes):
        p = _vp()
        self._check(self.L.hop_malloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, dptr):
        self._check(self.L.hop_free(self.h, _vp(dptr)))

    def pinned_array(self, shape, dtype=np.float32):
        """numpy array backed by pinned host memory (cudaHostAlloc); kept alive by the context."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _vp()
        self._check(self.L.hop_host_alloc(self.h, max(n, 1), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        if not hasattr(self, "_pinned"):
            self._pinned = []
        self._pinned.append((p.value, buf))
        return arr

    def h2d(self, dptr, host_array):
        a = np.ascontiguousarray(host_array)
        self._check(self.L.hop_memcpy_h2d(self.h, _vp(dptr), _ptr(a), a.nbytes))
        return a  # caller keeps it alive until sync

    def d2h(self, host_array, dptr):
        assert host_array.flags["C_CONTIGUOUS"]
        self._check(self.L.hop_memcpy_d2h(self.h, _ptr(host_array), _vp(dptr), host_array.nbytes))

    # ---- clouds ----
    def upload_cloud(self, xyz, nrm=None, prob=None):
        xyz = _f32(xyz, 3)
        nrm = None if nrm is None else _f32(nrm, 3)
        prob = None if prob is None else _f32(prob)
        h = _vp()
        self._check(self.L.hop_cloud_upload(self.h, _ptr(xyz), _ptr(nrm), _ptr(prob), len(xyz), C.byref(h)))
        return Cloud(self, h, len(xyz))

    def icp_params(self, **kw):
        p = IcpParams()
        self.L.hop_default_icp_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def lcp_params(self, **kw):
        p = LcpParams()
        self.L.hop_default_lcp_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    # ---- K4 / K5 / winners, host buffers (the reference-facing calls) ----
    def icp_refine(self, scene, model, poses, params=None):
        """poses (H,4,4) model2scene -> refined (H,4,4), iterations (H,), converged (H,)."""
        params = params or self.icp_params()
        flat = poses_to_colmajor(poses)
        H = len(flat)
        iters = np.zeros(H, np.int32)
        conv = np.zeros(H, np.int32)
        self._check(self.L.hop_icp_refine(self.h, scene.handle, model.handle, _ptr(flat), H, C.byref(params), _ptr(iters), _ptr(conv)))
        return colmajor_to_poses(flat), iters, conv

    def lcp_score(self, scene, model, poses, params=None, use_weights=False):
        params = params or self.lcp_params()
        flat = poses_to_colmajor(poses)
        H = len(flat)
        scores = np.zeros(H, np.float32)
        self._check(self.L.hop_lcp_score(self.h, scene.handle, model.handle, _ptr(flat), H, C.byref(params), int(use_weights), _ptr(scores)))
        return scores

    def verify_lcp(self, P_centered, Q_centered, bases, quads, quad_trial, centroid_P, centroid_Q, delta):
        """K3 over all congruent quadrilaterals of a frame.  Returns per-quad (poses (M,4,4), lcp, valid) and the
        emitted hypothesis list (hyp_poses (n,4,4), hyp_lcp) in (trial, quad) order."""
        Q = _f32(Q_centered, 3)
        bases = np.ascontiguousarray(bases, np.int32).reshape(-1, 4)
        quads = np.ascontiguousarray(quads, np.int32).reshape(-1, 4)
        qt = np.ascontiguousarray(quad_trial, np.int32)
        M = len(quads)
        cP, cQ = _f32(centroid_P), _f32(centroid_Q)
        poses = np.zeros((M, 16), np.float32); lcp = np.zeros(M, np.float32); valid = np.zeros(M, np.int32)
        hp = np.zeros((M, 16), np.float32); hl = np.zeros(M, np.float32); n = C.c_int32(0)
        self._check(self.L.hop_verify_lcp(self.h, P_centered.handle, _ptr(Q), len(Q), _ptr(bases), len(bases), _ptr(quads), _ptr(qt), M,
                                          _ptr(cP), _ptr(cQ), delta, _ptr(poses), _ptr(lcp), _ptr(valid), _ptr(hp), _ptr(hl), C.byref(n)))
        return colmajor_to_poses(poses), lcp, valid, colmajor_to_poses(hp[: n.value]), hl[: n.value]

    def super4pcs_run(self, plan, capacity=20000):
        """Device part of a planned registration.  Returns (poses (n,4,4) model -> scene, lcp (n,)) in (trial, quad) order."""
        poses = np.empty((capacity, 16), np.float32)   # (only the first n rows are defined afterwards)
        lcp = np.empty(capacity, np.float32)
        n = C.c_int32(0)
        self._check(self.L.hop_super4pcs_run(self.h, plan.h, _ptr(poses), _ptr(lcp), capacity, C.byref(n)))
        k = min(int(n.value), capacity)
        return colmajor_to_poses(poses[:k]), lcp[:k].copy()

    def hand_overlap(self, finger, scene_hand, scene_noswivel, params, thetas, scene_normals=None):
        """K1: cost[s] = objFuncPSO(thetas[s]) for S joint angles (radians); returns (cost (S,) float64, arg-min index)."""
        th = np.ascontiguousarray(thetas, np.float64)
        cost = np.zeros(len(th), np.float64)
        best = C.c_int32(-1)
        self._check(self.L.hop_hand_overlap(self.h, finger.handle, scene_hand.handle, scene_normals.handle if scene_normals else None,
                                            scene_noswivel.handle, C.byref(params), _ptr(th), len(th), _ptr(cost), C.byref(best)))
        return cost, int(best.value)

    def frame_params(self, K=None, cam_in_handbase=None, handbase_in_cam=None, **kw):
        """hop_frame_params with the reference's defaults; K = (fx, fy, cx, cy); the two 4x4 matrices are numpy (row-major) arrays --
        handbase_in_cam is `cam_in_handbase.inverse()` as the caller's host code computes it (float32 numeric inverse)."""
        p = FrameParams()
        self.L.hop_default_frame_params(C.byref(p))
        if K is not None:
            p.fx, p.fy, p.cx, p.cy = [float(v) for v in K]
        if cam_in_handbase is not None:
            T = np.asarray(cam_in_handbase, np.float32).reshape(4, 4)
            Ti = np.asarray(handbase_in_cam, np.float32).reshape(4, 4) if handbase_in_cam is not None else np.linalg.inv(T.astype(np.float64)).astype(np.float32)
            p.cam_in_handbase[:] = T.T.reshape(-1).tolist()
            p.handbase_in_cam[:] = Ti.T.reshape(-1).tolist()
        for k, v in kw.items():
            if k in ("box_min", "box_max", "viewpoint"):
                getattr(p, k)[:] = [float(x) for x in v]
            else:
                setattr(p, k, v)
        return p

    def frame_to_scene(self, depth_mm, params, scene=None):
        """the frame's front end on the device: uint16 depth image [mm] -> object-segment Cloud (+ the 5 stage counts)"""
        d = np.ascontiguousarray(depth_mm, np.uint16)
        h, w = d.shape
        handle = _vp(scene.handle.value if scene is not None else None)
        counts = np.zeros(5, np.int32)
        self._check(self.L.hop_frame_to_scene(self.h, _ptr(d), w, h, C.byref(params), C.byref(handle), _ptr(counts)))
        n = int(self.L.hop_cloud_size(handle))
        if scene is not None:
            scene.n = n
            return scene, counts
        return Cloud(self, handle, n), counts

    def debug_lm_solve(self, sums, with_cycles=False):
        """K4's inner solver alone on n moment sets (n x 96 float32: the packed upper triangle of the 13x13 moment matrix + 5 unused).
        Returns (x [n, 6], nfev [n], status [n]); status -1 = translation unconstrained."""
        sums = np.ascontiguousarray(sums, np.float32).reshape(-1, 96)
        n = len(sums)
        x = np.zeros((n, 6), np.float32); nfev = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
        cyc = np.zeros(n, np.int64)
        self._check(self.L.hop_debug_lm_solve(self.h, sums.ctypes.data, n, x.ctypes.data, nfev.ctypes.data, st.ctypes.data, cyc.ctypes.data if with_cycles else None))
        return (x, nfev, st, cyc) if with_cycles else (x, nfev, st)

    def ppf_table(self, xyz, nrm):
        """the model's PPF table (computePPF.cpp:56-107) on the device: distinct keys (n, 4) int32, sorted"""
        xyz, nrm = _f32(xyz, 3), _f32(nrm, 3)
        nk = C.c_int32(0)
        self._check(self.L.hop_ppf_table_build(self.h, _ptr(xyz), _ptr(nrm), len(xyz), None, 0, C.byref(nk)))
        keys = np.zeros((max(nk.value, 1), 4), np.int32)
        self._check(self.L.hop_ppf_table_build(self.h, _ptr(xyz), _ptr(nrm), len(xyz), _ptr(keys), nk.value, C.byref(nk)))
        return keys[: nk.value]

    def frame_organized(self, depth_mm, params, max_depth_change_factor=0.02, normal_smoothing_size=10.0):
        """depth image -> the frame's valid pixels (raster order) with PCL's integral-image normals (hop_frame_organized)"""
        depth_mm = np.ascontiguousarray(depth_mm, np.uint16)
        h, w = depth_mm.shape
        out = _vp()
        self._check(self.L.hop_frame_organized(self.h, _ptr(depth_mm), w, h, C.byref(params), max_depth_change_factor, normal_smoothing_size, C.byref(out)))
        return Cloud(self, out, int(self.L.hop_cloud_size(out)))

    def cluster_poses(self, poses, scores, angle_diff_deg, dist_diff, symmetry_deg=(360.0, 360.0, 360.0), ids=None):
        """PoseEstimator::clusterPoses with the comparisons on the device (hop_cluster_poses_gpu): the same keep list as the
        host function `cluster_poses`, in cluster order."""
        flat = poses_to_colmajor(poses)
        sc = _f32(scores)
        n = len(flat)
        idv = None if ids is None else np.ascontiguousarray(ids, np.int32)
        sym = np.ascontiguousarray(symmetry_deg, np.float32)
        keep = np.zeros(max(n, 1), np.int32)
        nk = C.c_int32(0)
        self._check(self.L.hop_cluster_poses_gpu(self.h, _ptr(flat), _ptr(sc), _ptr(idv), n, angle_diff_deg, dist_diff, _ptr(sym), _ptr(keep), C.byref(nk)))
        return keep[: nk.value].copy()

    def upload_mesh(self, V, F):
        """SDFchecker::registerMesh: V (nv,3) float vertices, F (nf,3) int faces"""
        V = _f32(V, 3)
        F = np.ascontiguousarray(F, np.int32).reshape(-1, 3)
        h = _vp()
        self._check(self.L.hop_mesh_upload(self.h, _ptr(V), len(V), _ptr(F), len(F), C.byref(h)))
        return Mesh(self, h, len(V), len(F))

    def sdf_query(self, mesh, pts, point_transforms=None, want_faces=False):
        """signed distance (igl pseudonormal rules) of pts under each of H point transforms (None = one identity placement):
        dict S (H,n), I (H,n) closest faces (when asked), min, max, n_inside (H,)"""
        pts = _f32(pts, 3)
        n = len(pts)
        flat = None if point_transforms is None else poses_to_colmajor(point_transforms)
        H = 1 if flat is None else len(flat)
        S = np.empty((H, n), np.float32)
        I = np.empty((H, n), np.int32) if want_faces else None
        mn, mx, cnt = np.empty(H, np.float32), np.empty(H, np.float32), np.empty(H, np.int32)
        self._check(self.L.hop_sdf_query(self.h, mesh.handle, _ptr(pts), n, _ptr(flat), H, _ptr(S), _ptr(I), _ptr(mn), _ptr(mx), _ptr(cnt)))
        return {"S": S, "I": I, "min": mn, "max": mx, "n_inside": cnt}

    def collision_params(self, d):
        p = CollisionParams()
        p.cam2handbase[:] = np.asarray(d["cam2handbase"], np.float32).T.reshape(-1).tolist()   # column-major
        p.model_center[:] = [float(v) for v in d["model_center"]]
        for k in ("ob_diameter", "collision_dist", "inside_ob_dist", "non_touch_dist", "collision_finger_dist", "collision_finger_volume_ratio"):
            setattr(p, k, float(d[k]))
        p.finger_status[:] = [int(v) for v in d["finger_status"]]
        return p

    @staticmethod
    def _handles4(objs):
        arr = (_vp * 4)()
        for k in range(4):
            o = objs[k] if objs is not None and k < len(objs) else None
            arr[k] = o.handle if o is not None else None
        return arr

    def reject_by_collision(self, object_mesh, finger_meshes, finger_clouds, scene_without_hand, hand_cloud, model, poses, params):
        """PoseEstimator::rejectByCollisionOrNonTouching for all poses: keep (H,) int32, reason (H,), diag (H,10)"""
        flat = poses_to_colmajor(poses)
        H = len(flat)
        keep, reason, diag = np.zeros(H, np.int32), np.zeros(H, np.int32), np.zeros((H, 10), np.float32)
        fm, fc = self._handles4(finger_meshes), self._handles4(finger_clouds)
        p = params if isinstance(params, CollisionParams) else self.collision_params(params)
        self._check(self.L.hop_reject_by_collision(self.h, object_mesh.handle, fm, fc, scene_without_hand.handle if scene_without_hand else None,
                                                   hand_cloud.handle if hand_cloud else None, model.handle if model else None, _ptr(flat), H,
                                                   C.byref(p), _ptr(keep), _ptr(reason), _ptr(diag)))
        return keep, reason, diag

    def reject_by_collision_dev(self, object_mesh, finger_meshes, finger_clouds, scene_without_hand, hand_cloud, model, d_poses, H, params,
                                d_keep, d_reason=None, d_diag=None):
        fm, fc = self._handles4(finger_meshes), self._handles4(finger_clouds)
        self._keepalive = (fm, fc, params)
        self._check(self.L.hop_reject_by_collision_dev(self.h, object_mesh.handle, fm, fc, scene_without_hand.handle if scene_without_hand else None,
                                                       hand_cloud.handle if hand_cloud else None, model.handle if model else None, d_poses, H,
                                                       C.byref(params), d_keep, d_reason, d_diag))

    TRIAL_HEIGHTS = (-0.03, -0.025, -0.02, -0.015, -0.01, -0.005, 0, 0.005, 0.01, 0.015, 0.02, 0.025, 0.03)   # Hand.cpp:1011

    def adjust_hand_height(self, hand_cloud, scene_handbase, heights=None):
        """HandT42::adjustHandHeight: (match counts per trial height, index of the chosen height or -1)"""
        hs = np.ascontiguousarray(self.TRIAL_HEIGHTS if heights is None else heights, np.float32)
        counts, best = np.zeros(len(hs), np.int32), C.c_int32(-1)
        self._check(self.L.hop_adjust_hand_height(self.h, hand_cloud.handle, scene_handbase.handle, _ptr(hs), len(hs), _ptr(counts), C.byref(best)))
        return counts, int(best.value)

    def hand_removal_params(self, handbase_in_cam, finger_1_2_in_handbase, finger_2_2_in_handbase, min_z, near_hand_dist):
        p = HandRemovalParams()
        hic = np.asarray(handbase_in_cam, np.float32)
        for name, M in (("cam_in_handbase", np.linalg.inv(hic)), ("handbase_in_cam", hic),
                        ("handbase_in_finger_1_2", np.linalg.inv(np.asarray(finger_1_2_in_handbase, np.float32))),
                        ("handbase_in_finger_2_2", np.linalg.inv(np.asarray(finger_2_2_in_handbase, np.float32)))):
            getattr(p, name)[:] = np.asarray(M, np.float32).T.reshape(-1).tolist()
        p.min_z = float(min_z)
        p.dist_thres_sq = float(np.float32(near_hand_dist) * np.float32(near_hand_dist))
        return p

    def remove_hand_points(self, scene, links, link_kind, params):
        """HandT42::removeSurroundingPointsAndAssignProbability: scene (Cloud, camera frame) -> new Cloud with confidences;
        links: Clouds in the hand-base frame in std::map order of their names, link_kind: 0 / 1 (proximal fingers) / 2 (base, swivels)"""
        arr = (_vp * max(len(links), 1))()
        for k, c in enumerate(links):
            arr[k] = c.handle if c is not None else None
        kinds = np.ascontiguousarray(link_kind, np.int32)
        h = _vp()
        self._check(self.L.hop_remove_hand_points(self.h, scene.handle, arr, _ptr(kinds), len(links), C.byref(params), C.byref(h)))
        return Cloud(self, h, int(self.L.hop_cloud_size(h)))

    def render_params(self, **kw):
        p = RenderParams()
        self.L.hop_default_render_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def render_scene(self, params, depth_m, hand_V=None, hand_F=None):
        """Renderer set-up of rejectByRender: depth_m (height,width) float metres, hand meshes in the camera frame"""
        depth_m = np.ascontiguousarray(depth_m, np.float32)
        assert depth_m.shape == (params.height, params.width)
        hV = None if hand_V is None or len(hand_V) == 0 else _f32(hand_V, 3)
        hF = None if hand_F is None or len(hand_F) == 0 else np.ascontiguousarray(hand_F, np.int32).reshape(-1, 3)
        h = _vp()
        self._check(self.L.hop_render_scene_create(self.h, C.byref(params), _ptr(depth_m), _ptr(hV), 0 if hV is None else len(hV), _ptr(hF),
                                                   0 if hF is None else len(hF), C.byref(h)))
        return RenderScene(self, h, params)

    def render_depth(self, scene, obj_V, obj_F, pose):
        """one simulated depth image (metres) + the object mask"""
        V, F = _f32(obj_V, 3), np.ascontiguousarray(obj_F, np.int32).reshape(-1, 3)
        flat = poses_to_colmajor(np.asarray(pose, np.float32).reshape(1, 4, 4))
        depth = np.empty((scene.params.height, scene.params.width), np.float32)
        mask = np.empty((scene.params.height, scene.params.width), np.uint8)
        self._check(self.L.hop_render_depth(self.h, scene.handle, _ptr(V), len(V), _ptr(F), len(F), _ptr(flat), _ptr(depth), _ptr(mask)))
        return depth, mask

    def reject_by_render(self, scene, obj_V, obj_F, poses):
        """PoseEstimator::rejectByRender: (wrong_ratio (H,), kept hypothesis indices in the reference's output order)"""
        V, F = _f32(obj_V, 3), np.ascontiguousarray(obj_F, np.int32).reshape(-1, 3)
        flat = poses_to_colmajor(poses)
        H = len(flat)
        wr, order, nk = np.zeros(H, np.float32), np.zeros(max(H, 1), np.int32), C.c_int32(0)
        self._check(self.L.hop_reject_by_render(self.h, scene.handle, _ptr(V), len(V), _ptr(F), len(F), _ptr(flat), H, _ptr(wr), _ptr(order), C.byref(nk)))
        return wr, order[: nk.value].copy()

    def select_topk(self, poses, scores, K, id_offset=0, frame=0):
        flat = poses_to_colmajor(poses)
        scores = _f32(scores)
        out = np.zeros(K, POSE_REC_DTYPE)
        self._check(self.L.hop_select_topk(self.h, _ptr(flat), _ptr(scores), len(flat), K, id_offset, frame, _ptr(out)))
        return out

    def refine_score_select(self, scene, model_icp, poses, K=1, model_lcp=None, icp_params=None, lcp_params=None, use_weights=False):
        """refineByICP + selectBest in one visit to the device (hop_refine_score_select): returns refined poses (H,4,4), scores,
        iterations, converged, winners (K records, score descending)."""
        icp_params = icp_params or self.icp_params()
        lcp_params = lcp_params or self.lcp_params()
        flat = poses_to_colmajor(poses)
        H = len(flat)
        scores, iters, conv = np.zeros(H, np.float32), np.zeros(H, np.int32), np.zeros(H, np.int32)
        win = np.zeros(K, POSE_REC_DTYPE)
        self._check(self.L.hop_refine_score_select(self.h, scene.handle, model_icp.handle, model_lcp.handle if model_lcp else None, _ptr(flat), H,
                                                   C.byref(icp_params), C.byref(lcp_params), int(use_weights), K, _ptr(flat), _ptr(scores),
                                                   _ptr(iters), _ptr(conv), _ptr(win) if K else None))
        return colmajor_to_poses(flat), scores, iters, conv, win

    # ---- the one collective: all-gather of winner records through the C ABI (NCCL bound at run time) ----
    def comm_init(self, unique_id, rank, world):
        """unique_id: the 128 bytes rank 0 got from comm_unique_id() and handed to every rank"""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.L.hop_comm_init(self.h, buf, int(rank), int(world)))

    def gather_winners(self, local):
        """host records (numpy POSE_REC_DTYPE, K) -> world x K records, rank-major"""
        local = np.ascontiguousarray(local)
        r, w = C.c_int(0), C.c_int(1)
        self._check(self.L.hop_comm_rank(self.h, C.byref(r), C.byref(w)))
        out = np.zeros(w.value * len(local), POSE_REC_DTYPE)
        self._check(self.L.hop_gather_winners(self.h, _ptr(local), len(local), _ptr(out)))
        return out

    def gather_winners_dev(self, d_send, K, d_recv, overlap=False):
        self._check(self.L.hop_gather_winners_dev(self.h, _vp(d_send), int(K), _vp(d_recv), int(overlap)))

    def gather_wait(self, block_host=False):
        self._check(self.L.hop_gather_wait(self.h, int(block_host)))

    # ---- device-pointer variants (inputs already resident in HBM) ----
    def refine_score_select_dev(self, scene, model_icp, d_poses, H, icp_params, lcp_params, K, d_scores, d_winners, model_lcp=None,
                                d_iters=None, d_conv=None, use_weights=False, id_offset=0, frame=0):
        self._check(self.L.hop_refine_score_select_dev(self.h, scene.handle, model_icp.handle, model_lcp.handle if model_lcp else None,
                                                       _vp(d_poses), H, C.byref(icp_params), C.byref(lcp_params), int(use_weights), K,
                                                       id_offset, frame, _vp(d_iters) if d_iters else None, _vp(d_conv) if d_conv else None,
                                                       _vp(d_scores), _vp(d_winners) if d_winners else None))

    def icp_refine_dev(self, scene, model, d_poses, H, params, d_iters=None, d_conv=None):
        self._check(self.L.hop_icp_refine_dev(self.h, scene.handle, model.handle, _vp(d_poses), H, C.byref(params),
                                              _vp(d_iters) if d_iters else None, _vp(d_conv) if d_conv else None))

    def lcp_score_dev(self, scene, model, d_poses, H, params, d_scores, use_weights=False):
        self._check(self.L.hop_lcp_score_dev(self.h, scene.handle, model.handle, _vp(d_poses), H, C.byref(params),
                                             int(use_weights), _vp(d_scores)))

    def select_topk_dev(self, d_poses, d_scores, H, K, d_out, id_offset=0, frame=0):
        self._check(self.L.hop_select_topk_dev(self.h, _vp(d_poses), _vp(d_scores), H, K, id_offset, frame, _vp(d_out)))
