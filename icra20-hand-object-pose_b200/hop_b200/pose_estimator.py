"""Host mirror of the reference's PoseEstimator stage object for the hot path.

Same method names, argument meaning and error behaviour as src/perception/include/PoseEstimator.h:11-49 /
src/perception/src/PoseEstimator.cpp for the calls on the north-star path:
    refineByICP()   PoseEstimator.cpp:235-275  -> hop_icp_refine  (K4)
    selectBest()    PoseEstimator.cpp:465-502  -> hop_lcp_score   (K5) + arg-max
PCL is absent, so clouds are (xyz, normal[, confidence]) numpy arrays instead of pcl::PointCloud<PointT>::Ptr.
The native C++ mirror of the same class lives in ../host/ (used by main_realdata_auto).
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class PoseHypo:
    """class PoseHypo (src/perception/include/PoseHypo.h:7-27)."""
    _pose: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    _id: int = -1
    _lcp_score: float = 0.0
    _wrong_ratio: float = 1.0


class PoseEstimator:
    def __init__(self, ctx, cfg=None):
        """cfg: dict with the config_autodataset.yaml keys used on this path (icp_dist_thres, icp_angle_thres, lcp.*)."""
        self.ctx = ctx
        cfg = cfg or {}
        self.icp_dist_thres = float(cfg.get("icp_dist_thres", 0.01))
        self.icp_angle_thres = float(cfg.get("icp_angle_thres", 45))
        lcp = cfg.get("lcp", {})
        self.lcp_dist = float(lcp.get("dist", 0.001))
        self.lcp_normal_angle = float(lcp.get("normal_angle", 10))
        self.max_icp_candidates = 100  # PoseEstimator.cpp:241
        self.gpu_cluster_min = 2048    # clusterPoses: device version from this many hypotheses up (same result)
        sym = cfg.get("object_symmetry", {}).get(cfg.get("model_name", ""), {})
        self.object_symmetry = (float(sym.get("x", 360)), float(sym.get("y", 360)), float(sym.get("z", 360)))
        self._pose_hypos = []
        self._scene = self._model = self._model001 = None
        self._scene_host = self._model_host = None     # host copies: the Super4PCS planner runs on the host

    # -- setCurScene (PoseEstimator.cpp:36-45): keeps the high-confidence scene (confidence >= thres)
    def setCurScene(self, scene_xyz, scene_nrm, confidence=None, high_confidence_thres=0.8):
        scene_xyz = np.asarray(scene_xyz, np.float32)
        scene_nrm = np.asarray(scene_nrm, np.float32)
        if confidence is not None:
            keep = np.asarray(confidence) >= high_confidence_thres
            scene_xyz, scene_nrm, confidence = scene_xyz[keep], scene_nrm[keep], np.asarray(confidence, np.float32)[keep]
        self._scene_host = (scene_xyz, scene_nrm, confidence)
        if self._scene is None:
            self._scene = self.ctx.upload_cloud(scene_xyz, scene_nrm, confidence)
        else:
            self._scene.update(scene_xyz, scene_nrm, confidence)

    def setModel(self, model_xyz, model_nrm, model001_xyz=None, model001_nrm=None):
        """_model (5 mm, ICP / Super4PCS) and _model001 (1 mm, scoring) (main_realdata_auto.cpp:33-38)."""
        self._model = self.ctx.upload_cloud(model_xyz, model_nrm).hint_static()
        self._model_host = (np.asarray(model_xyz, np.float32), np.asarray(model_nrm, np.float32))
        if model001_xyz is None:
            self._model001 = self._model
        else:
            self._model001 = self.ctx.upload_cloud(model001_xyz, model001_nrm).hint_static()
            self._model001_host = np.asarray(model001_xyz, np.float32)

    def setPoseHypos(self, poses, scores=None):
        poses = np.asarray(poses, np.float32).reshape(-1, 4, 4)
        scores = np.zeros(len(poses), np.float32) if scores is None else np.asarray(scores, np.float32)
        self._pose_hypos = [PoseHypo(poses[i].copy(), i, float(scores[i])) for i in range(len(poses))]

    def runSuper4pcs(self, ppfs, **options):
        """PoseEstimator::runSuper4pcs (PoseEstimator.cpp:62-100): source = _model, target = _scene_high_confidence; every
        congruent quadrilateral with LCP > 0 becomes PoseHypo(pose, i, lcp).  ppfs: (n,4) int keys of the model's PPF table
        (the reference passes the whole std::map; only key membership is ever used).  Returns False when nothing was found
        (the caller prints "No pose found" and exits, main_realdata_auto.cpp:189-196)."""
        from . import capi
        if self._scene_host is None or self._model_host is None:
            raise RuntimeError("runSuper4pcs needs setCurScene() and setModel() first")
        sx, sn, sc = self._scene_host
        mx, mn = self._model_host
        plan = capi.S4pcsPlan(sx, sn, sc, mx, mn, ppfs, capi.s4pcs_options(**options))
        try:
            poses, lcp = self.ctx.super4pcs_run(plan)
        finally:
            plan.close()
        self._pose_hypos = [PoseHypo(poses[i].copy(), i, float(lcp[i])) for i in range(len(poses))]
        return len(self._pose_hypos) > 0

    def clusterPoses(self, angle_diff, dist_diff, assign_id=False):
        """PoseEstimator::clusterPoses(angle_diff [deg], dist_diff [m], assign_id) (PoseEstimator.cpp:106-233): greedy
        suppression in (lcp desc, id asc) order with the object's symmetry (object_symmetry.<model_name>.{x,y,z})."""
        from . import capi
        if not self._pose_hypos:
            raise IndexError("clusterPoses on an empty hypothesis list (the reference reads hypo_tmp[0])")
        poses = np.stack([h._pose for h in self._pose_hypos])
        scores = np.array([h._lcp_score for h in self._pose_hypos], np.float32)
        ids = np.array([h._id for h in self._pose_hypos], np.int32)
        # the device version decides identically; it pays from a few thousand hypotheses up
        if len(poses) >= self.gpu_cluster_min:
            keep = self.ctx.cluster_poses(poses, scores, angle_diff, dist_diff, self.object_symmetry, ids)
        else:
            keep = capi.cluster_poses(poses, scores, angle_diff, dist_diff, self.object_symmetry, ids)
        self._pose_hypos = [self._pose_hypos[k] for k in keep]
        if assign_id:
            for i, h in enumerate(self._pose_hypos):
                h._id = i

    def refineByICP(self):
        """Keeps the first min(N,100) hypotheses and replaces each pose by T_icp^-1 * pose (PoseEstimator.cpp:235-275)."""
        keep = self._pose_hypos[: min(len(self._pose_hypos), self.max_icp_candidates)]
        if not keep:
            self._pose_hypos = []
            return
        poses = np.stack([h._pose for h in keep])
        params = self.ctx.icp_params(max_iter=10, angle_deg=self.icp_angle_thres, max_dist=self.icp_dist_thres)
        # the scene grid selectBest's computeLCP needs depends on the frame only: built on the second stream while the ICP runs
        self._scene.prepare_lcp_scene(self.ctx.lcp_params(dist=self.lcp_dist, angle_deg=self.lcp_normal_angle))
        refined, _, _ = self.ctx.icp_refine(self._scene, self._model, poses, params)
        for h, p in zip(keep, refined):
            h._pose = p
        self._pose_hypos = keep

    # -- physics pruning ---------------------------------------------------------------------------
    def registerMesh(self, V, F, name="object"):
        """PoseEstimator::registerMesh / registerHandMesh (PoseEstimator.cpp:506-521): name = "object" or a finger link
        ("finger_1_1", "finger_1_2", "finger_2_1", "finger_2_2", vertices already in the hand-base frame)."""
        if not hasattr(self, "_meshes"):
            self._meshes = {}
        if name in self._meshes:
            self._meshes[name].free()
        self._meshes[name] = self.ctx.upload_mesh(V, F)
        if not hasattr(self, "_mesh_host"):
            self._mesh_host = {}
        self._mesh_host[name] = (np.asarray(V, np.float32), np.asarray(F, np.int32))

    FINGERS = ("finger_1_1", "finger_1_2", "finger_2_1", "finger_2_2")

    def rejectByCollisionOrNonTouching(self, hand, cfg=None):
        """PoseEstimator::rejectByCollisionOrNonTouching(HandT42*) (PoseEstimator.cpp:524-735).  `hand`: dict with
        component_status {name: bool}, finger_clouds {name: xyz in the hand-base frame}, hand_cloud (xyz, hand-base frame),
        handbase_in_cam (4x4), cloud_withouthand (xyz in the hand-base frame, 5 mm voxel-sampled).  Survivors keep their order."""
        cfg = cfg or {}
        if not cfg.get("pose_estimator_use_physics", True):
            print("Not using physics")
            return
        if not self._pose_hypos:
            return
        mx = self._model_host[0]
        m001 = mx if getattr(self, "_model001_host", None) is None else self._model001_host
        ext = m001.max(0) - m001.min(0)
        smallest, diam = float(ext.min()), float(np.linalg.norm(ext))
        status = [bool(hand["component_status"].get(n, False)) for n in self.FINGERS]
        params = dict(cam2handbase=np.linalg.inv(np.asarray(hand["handbase_in_cam"], np.float32)), model_center=m001.mean(0),
                      ob_diameter=diam, collision_dist=min(-smallest * float(cfg.get("collision_thres", 0.4)), -0.007),
                      inside_ob_dist=min(-smallest / 5, -0.01), non_touch_dist=float(cfg.get("non_touch_dist", 0.01)),
                      collision_finger_dist=-float(cfg.get("collision_finger_dist", 0.012)),
                      collision_finger_volume_ratio=float(cfg.get("collision_finger_volume_ratio", 0.25)), finger_status=status)
        meshes = getattr(self, "_meshes", {})
        fclouds = [self.ctx.upload_cloud(hand["finger_clouds"][n]) if status[k] and n in hand["finger_clouds"] else None
                   for k, n in enumerate(self.FINGERS)]
        scene = self.ctx.upload_cloud(hand["cloud_withouthand"]) if len(hand.get("cloud_withouthand", ())) else None
        hcloud = self.ctx.upload_cloud(hand["hand_cloud"]) if len(hand.get("hand_cloud", ())) else None
        poses = np.stack([h._pose for h in self._pose_hypos])
        keep, reason, _ = self.ctx.reject_by_collision(meshes["object"], [meshes.get(n) for n in self.FINGERS], fclouds, scene, hcloud,
                                                       self._model, poses, params)
        for c in fclouds + [scene, hcloud]:
            if c is not None:
                c.free()
        self._pose_hypos = [h for h, k in zip(self._pose_hypos, keep) if k]
        self._reject_reason = reason

    def rejectByRender(self, projection_thres, hand, depth_meters, cam_K, cfg=None):
        """PoseEstimator::rejectByRender(projection_thres, HandT42*) (PoseEstimator.cpp:345-463): renders the hand + the object under
        every hypothesis, compares with the real depth image and keeps the max(render_keep_hypo * N, 10) hypotheses with the lowest
        wrong ratio, in ascending order (projection_thres is unused by the reference too).  `hand`: dict with component_status
        {name: bool}, meshes {name: (V in the link frame, F)}, tf_in_base {name: 4x4 getTFHandBase}, handbase_in_cam; the object
        mesh is the one given to registerMesh(V, F, "object") (kept on the host for the renderer)."""
        cfg = cfg or {}
        print(f"before projection check, #hypo={len(self._pose_hypos)}")
        if not self._pose_hypos:
            return
        K = np.asarray(cam_K, np.float32).reshape(3, 3)
        h, w = np.asarray(depth_meters).shape
        params = self.ctx.render_params(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]), width=int(w), height=int(h),
                                        roi_weight=float(cfg.get("render_roi_weight", 2.0)), keep_ratio=float(cfg.get("render_keep_hypo", 0.3)))
        hV, hF, off = [], [], 0
        for name, (V, F) in sorted(hand.get("meshes", {}).items()):          # std::map order
            if not hand["component_status"].get(name, False):
                continue
            T = np.asarray(hand["handbase_in_cam"], np.float32) @ np.asarray(hand["tf_in_base"][name], np.float32)
            hV.append((np.asarray(V, np.float32) @ T[:3, :3].T + T[:3, 3]).astype(np.float32))
            hF.append(np.asarray(F, np.int32) + off)
            off += len(V)
        scene = self.ctx.render_scene(params, depth_meters, np.concatenate(hV) if hV else None, np.concatenate(hF) if hF else None)
        oV, oF = self._mesh_host["object"]
        poses = np.stack([hy._pose for hy in self._pose_hypos])
        wr, order = self.ctx.reject_by_render(scene, oV, oF, poses)
        scene.free()
        for hy, r in zip(self._pose_hypos, wr):
            hy._wrong_ratio = float(r)
        self._pose_hypos = [self._pose_hypos[k] for k in order]
        print(f"after projection check, #hypo={len(self._pose_hypos)}")

    def selectBest(self):
        """Scores every hypothesis with computeLCP(1 mm model, lcp.dist, lcp.normal_angle, true,true,true) and returns
        the arg-max (first best, best_lcp starts at 0 -> hypothesis 0 when nothing scores) (PoseEstimator.cpp:465-502)."""
        if not self._pose_hypos:
            raise IndexError("selectBest on an empty hypothesis list (the reference dereferences _pose_hypos[0])")
        poses = np.stack([h._pose for h in self._pose_hypos])
        params = self.ctx.lcp_params(dist=self.lcp_dist, angle_deg=self.lcp_normal_angle)
        scores = self.ctx.lcp_score(self._scene, self._model001, poses, params)
        best, best_lcp = self._pose_hypos[0], 0.0
        for h, s in zip(self._pose_hypos, scores):
            h._lcp_score = float(s)
            if h._lcp_score > best_lcp:
                best, best_lcp = h, h._lcp_score
        return best
