"""Host mirror of the reference's hand-state search for one finger link.

    FingerProperty                 src/perception/src/Hand.cpp:182-250   (bounding box + per-z-bin extents of a link cloud)
    Hand::matchOneComponentPSO     src/perception/src/Hand.cpp:603-672   -> hop_hand_overlap (K1) over a dense angle grid

The reference minimises objFuncPSO with a 16-particle, 3-generation swarm (64 evaluations, Armadillo RNG).  Here the
whole admissible interval is evaluated at S evenly spaced angles in one launch and the arg-min is taken on the device:
a superset of anything the swarm can visit up to the grid pitch, deterministic, and with no per-thread copies of clouds
or trees.  Same name, argument meaning and failure behaviour (finger disabled when -cost <= least_match).
"""
import numpy as np

from .capi import FingerParams, HOP_MAX_FINGER_BINS


class FingerProperty:
    """FingerProperty(model, num_division) (Hand.cpp:184-236); all arithmetic in float32 like the reference."""

    def __init__(self, xyz, num_division=10):
        xyz = np.asarray(xyz, np.float32)
        assert 1 <= num_division <= HOP_MAX_FINGER_BINS
        self._num_division = num_division
        mn, mx = xyz.min(0), xyz.max(0)
        self._min_x, self._min_y, self._min_z = (np.float32(v) for v in mn)
        self._max_x, self._max_y, self._max_z = (np.float32(v) for v in mx)
        self._stride_z = np.float32((self._max_z - self._min_z) / np.float32(num_division))
        big = np.finfo(np.float32).max
        hist = np.empty((6, num_division), np.float32)
        hist[:3], hist[3:] = big, -big
        bins = self.getBinAlongZ(xyz[:, 2])
        changed = np.zeros(num_division, bool)
        for b in range(num_division):
            sel = xyz[bins == b]
            if len(sel):
                hist[:3, b], hist[3:, b] = sel.min(0), sel.max(0)
                changed[b] = True
        for i in range(num_division):                 # untouched bin: the next touched one (Hand.cpp:213-225)
            if not changed[i]:
                for j in range(i + 1, num_division):
                    if changed[j]:
                        hist[:, i] = hist[:, j]
                        changed[i] = True
                        break
        if not changed[-1]:
            for i in range(num_division - 2, -1, -1):
                if changed[i]:
                    hist[:, -1] = hist[:, i]
                    break
        self._hist_alongz = hist

    def getBinAlongZ(self, z):
        z = np.asarray(z, np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            b = (np.maximum(z - self._min_z, np.float32(0)) / self._stride_z).astype(np.int64)  # truncation, like the int cast
        return np.clip(b, 0, self._num_division - 1)


def finger_params(prop, scalars):
    """hop_finger_params from a FingerProperty and the ArgPasser / YAML scalars (dict, see synth.make_hand_case)."""
    p = FingerParams()
    m = np.asarray(scalars["model2handbase"], np.float32).reshape(4, 4)
    o = np.asarray(scalars.get("finger_out2parent", np.eye(4)), np.float32).reshape(4, 4)
    p.model2handbase[:] = m.T.reshape(-1).tolist()    # column-major
    p.finger_out2parent[:] = o.T.reshape(-1).tolist()
    p.tip1_local[:] = [float(v) for v in scalars["tip1_local"]]
    p.tip2_local[:] = [float(v) for v in scalars["tip2_local"]]
    for k in ("pair_tip1_y", "pair_tip2_y", "gripper_min_dist", "dist_thres", "normal_angle_deg", "outter_pt_dist", "outter_pt_dist_weight"):
        setattr(p, k, float(scalars[k]))
    for k in ("palm_side", "right_side", "check_normal", "max_outter_pts"):
        setattr(p, k, int(scalars[k]))
    p.num_division = prop._num_division
    p.min_z, p.stride_z = float(prop._min_z), float(prop._stride_z)
    hist = np.zeros(HOP_MAX_FINGER_BINS, np.float32)
    hist[: prop._num_division] = prop._hist_alongz[1]
    p.hist_min_y[:] = hist.tolist()
    return p


class HandMatcher:
    """The part of class Hand the hot path needs: per-link clouds + the scene of the current frame, on the device."""

    def __init__(self, ctx, n_states=4096):
        self.ctx, self.n_states = ctx, n_states
        self._clouds, self._props = {}, {}
        self._scene = self._lookup = self._noswivel = None
        self._tf_self, self._component_status = {}, {}
        self.objval = None

    def addComponent(self, name, xyz, nrm, num_division=10):
        self._clouds[name] = self.ctx.upload_cloud(xyz, nrm)
        self._props[name] = FingerProperty(xyz, num_division)

    def setCurScene(self, scene_xyz, scene_nrm, noswivel_xyz, noswivel_nrm=None, lookup_xyz=None, lookup_nrm=None):
        """scene = scene_hand_region_removed_noise (hand-base frame), noswivel = scene_remove_swivel, lookup = the cloud
        whose normals the reference reads with the neighbour index (Hand.cpp:94,326)."""
        for c in (self._scene, self._lookup, self._noswivel):
            if c is not None:
                c.free()
        self._scene = self.ctx.upload_cloud(scene_xyz, scene_nrm)
        self._noswivel = self.ctx.upload_cloud(noswivel_xyz, noswivel_nrm)
        self._lookup = None if lookup_xyz is None else self.ctx.upload_cloud(lookup_xyz, lookup_nrm)

    def matchOneComponentPSO(self, model_name, min_angle, max_angle, scalars, least_match=5.0):
        """min/max angle in degrees (Hand.cpp:603).  Returns success; on success _tf_self[model_name] holds Rx(angle)."""
        lo, hi = np.float32(min_angle) * np.pi / 180, np.float32(max_angle) * np.pi / 180  # `min_angle*M_PI/180`: float * double
        thetas = lo + (hi - lo) * (np.arange(self.n_states, dtype=np.float64) / max(self.n_states - 1, 1))
        params = finger_params(self._props[model_name], scalars)
        cost, best = self.ctx.hand_overlap(self._clouds[model_name], self._scene, self._noswivel, params, thetas, self._lookup)
        self.objval = float(cost[best])
        self.costs, self.thetas = cost, thetas
        if not (-self.objval > least_match):
            self._tf_self[model_name] = np.eye(4, dtype=np.float32)
            self._component_status[model_name] = False
            return False
        angle = np.float32(thetas[best])
        T = np.eye(4, dtype=np.float32)
        c, s = np.cos(angle), np.sin(angle)
        T[1, 1], T[1, 2], T[2, 1], T[2, 2] = c, -s, s, c
        self._tf_self[model_name] = T
        self._component_status[model_name] = True
        self.angle = float(angle)
        return True


def set_cur_scene(scene_hand_region, handbase_in_cam):
    """The cloud chain of Hand::setCurScene (Hand.cpp:279-334) on the device, from the cropped hand-region cloud (a capi.Cloud in
    the camera frame, with normals): VoxelGrid 3 mm -> into the hand-base frame -> RadiusOutlierRemoval(0.02, 30) ->
    RadiusOutlierRemoval(0.04, 100) -> StatisticalOutlierRemoval(20, 2) -> PassThrough x in [-0.25, -0.1].
    Returns the three clouds _pso_args holds (all in the hand-base frame): scene_hand_region (what the normals are looked up in),
    scene_hand_region_removed_noise (what the kd-tree indexes), scene_remove_swivel.  (handbaseICP, Hand.cpp:677-763, runs before
    this on the organised scene through hop_icp_refine and may have moved handbase_in_cam.)"""
    ds = scene_hand_region.voxel_grid(0.003)
    hb = ds.transform(np.linalg.inv(np.asarray(handbase_in_cam, np.float32)))
    r1 = hb.radius_outlier_removal(0.02, 30)
    r2 = r1.radius_outlier_removal(0.04, 100)
    clean = r2.statistical_outlier_removal(20, 2.0)
    noswivel = clean.pass_through("x", -0.25, -0.1)
    for c in (ds, r1, r2):
        c.free()
    return {"scene_hand_region": hb, "scene_hand_region_removed_noise": clean, "scene_remove_swivel": noswivel}


def euler_zyx(R):
    """Eigen's Matrix3f::eulerAngles(2, 1, 0) (first angle in [0, pi]), as Hand::handbaseICP reads the pitch (Hand.cpp:747-748)"""
    R = np.asarray(R, np.float32)
    a0 = np.float32(np.arctan2(R[1, 0], R[0, 0]))
    c2 = np.float32(np.hypot(R[2, 2], R[2, 1]))
    if a0 < 0:
        a0 = np.float32(a0 + np.float32(np.pi))
        a1 = np.float32(np.arctan2(-R[2, 0], -c2))
    else:
        a1 = np.float32(np.arctan2(-R[2, 0], c2))
    s1, c1 = np.float32(np.sin(a0)), np.float32(np.cos(a0))
    a2 = np.float32(np.arctan2(s1 * R[0, 2] - c1 * R[1, 2], c1 * R[1, 1] - s1 * R[0, 1]))
    return np.array([a0, a1, a2], np.float32)


def handbase_icp(ctx, scene_organized, base_link_cloud, handbase_in_cam, finger_1_1_in_parent, finger_2_1_in_parent):
    """Hand::handbaseICP (Hand.cpp:677-763) on the device: VoxelGrid 5 mm -> into the hand-base frame -> PassThrough x [-0.07, 0.03],
    z [-0.18, 0.01] -> finger connection parts removed -> runICP(scene -> base_link cloud, 50 iterations, 30 deg, 0.03 m) -> the
    reference's sanity gates.  scene_organized / base_link_cloud: capi.Cloud (camera frame / hand-base frame, with normals).
    Returns (handbase_in_cam after the correction, handbase matched?, cam2handbase_offset)."""
    hic = np.asarray(handbase_in_cam, np.float32)
    ds = scene_organized.voxel_grid(0.005)
    hb = ds.transform(np.linalg.inv(hic))
    px = hb.pass_through("x", -0.07, 0.03)
    pz = px.pass_through("z", -0.18, 0.01)
    f1, f2 = np.asarray(finger_1_1_in_parent, np.float32), np.asarray(finger_2_1_in_parent, np.float32)
    region = pz.handbase_region(float(f1[1, 3]), float(f1[2, 3]), float(f2[1, 3]), float(f2[2, 3]))
    offset = np.eye(4, dtype=np.float32)
    if region.n > 0 and base_link_cloud.n > 0:
        params = ctx.icp_params(max_iter=50, angle_deg=30.0, max_dist=0.03)
        # Utils::runICP(src = scene_handbase, tgt = handbase, T): hop_icp_refine refines poses of the TARGET (model -> scene),
        # i.e. pose <- T^-1 * pose; starting from the identity the returned pose is T^-1
        refined, _, _ = ctx.icp_refine(region, base_link_cloud, np.eye(4, dtype=np.float32)[None], params)
        offset = np.linalg.inv(refined[0].astype(np.float64)).astype(np.float32)
    for c in (ds, hb, px, pz, region):
        c.free()
    translation = float(np.linalg.norm(offset[:3, 3]))
    if translation >= 0.05:
        offset = np.eye(4, dtype=np.float32)
    # Utils::rotationGeodesicDistance(I, R) (Utils.cpp:29-32): acos((trace - 1) / 2), clamped
    cosv = float(np.clip((np.trace(offset[:3, :3]) - 1.0) / 2.0, -1.0, 1.0))
    rot_diff = np.degrees(np.arccos(cosv))
    pitch = float(euler_zyx(offset[:3, :3])[1])
    pitch = min(abs(pitch), abs(float(np.float32(np.pi)) - pitch))
    pitch = min(abs(pitch), abs(float(np.float32(np.pi)) + pitch))
    if rot_diff >= 10 or abs(pitch) >= 10 / 180.0 * np.pi:
        offset = np.eye(4, dtype=np.float32)
    matched = not np.array_equal(offset, np.eye(4, dtype=np.float32))
    return (hic.astype(np.float64) @ np.linalg.inv(offset.astype(np.float64))).astype(np.float32), matched, offset
