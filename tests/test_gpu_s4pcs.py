"""GPU parity tests of the Super4PCS device stages (K2a pair extraction, K2b congruent-set search, K3 verification) run
through hop_super4pcs_run, against the reference's OWN compiled matcher (oracle/_ref).

Bars (index/integer work): per trial, the SET of extracted pairs, the SET of congruent quadrilaterals and the multiset
of emitted hypotheses (pose bits + LCP) must equal the reference's; the number of executed trials must be the same.
Inside a trial the reference lists pairs in its octree-traversal order, the device in (i, j) order: order is not compared
(the reference's own OpenMP build does not keep it either, congruentSetExplorationBase.hpp:324-333)."""
import numpy as np
import pytest

from hop_b200 import capi, synth
from oracle import cpu_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(O.ref() is None or not hasattr(O.ref(), "hop_ref_s4pcs_get_trials"), reason="oracle/_ref (compiled OpenGR) not built")]


def _setof(a):
    return set(map(tuple, np.asarray(a).tolist()))


def _compare(ctx, name, seed, nq, ns, **opts):
    m, mn = synth.make_model(name, nq, seed=1)
    keys = O.ref_ppf_keys(m[:400], mn[:400])
    s, sn, conf, gt = synth.make_scene(name, ns, seed=seed)
    r = O.ref_super4pcs(s, sn, conf, m, mn, keys, **opts)
    plan = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options(keep_intermediates=1, **opts))
    poses, lcp = ctx.super4pcs_run(plan)
    ranges, pairs, quads = plan.intermediates()
    T = len(r["base_ok"])
    assert plan.sizes()["trials_executed"] == T
    n_pairs = n_quads = 0
    for t in range(T):
        pr = r["pair_ranges"][t]
        if not r["base_ok"][t]:
            assert ranges[t, 1] == ranges[t, 0] and ranges[t, 3] == ranges[t, 2]
            continue
        for k in (0, 1):                                            # K2a: both pair sets of the trial
            ref_pairs = r["pairs"][pr[2 * k]:pr[2 * k + 1]]
            got_pairs = pairs[ranges[t, 2 * k]:ranges[t, 2 * k + 1]]
            assert len(got_pairs) == len(ref_pairs) and _setof(got_pairs) == _setof(ref_pairs), (t, k)
            n_pairs += len(ref_pairs)
        tr = r["trials"][t]                                         # K2b: the congruent set
        ref_quads = r["quads"][tr[5]:tr[6]] if tr[4] else np.zeros((0, 4), np.int32)
        got_quads = quads[ranges[t, 4]:ranges[t, 5]]
        assert len(got_quads) == len(ref_quads) and _setof(got_quads) == _setof(ref_quads), t
        n_quads += len(ref_quads)
    # K3 + compaction: the emitted hypotheses, as a multiset of (pose bits, lcp)
    assert len(poses) == len(r["poses"])
    key = lambda P, L: sorted(zip(map(bytes, np.ascontiguousarray(np.round(P.reshape(len(P), -1), 6))), L.tolist()))
    got_sorted = np.array(sorted(np.concatenate([poses.reshape(len(poses), -1), lcp[:, None]], 1).tolist()))
    ref_sorted = np.array(sorted(np.concatenate([r["poses"].reshape(len(poses), -1), r["lcp"][:, None]], 1).tolist()))
    assert np.array_equal(got_sorted[:, -1], ref_sorted[:, -1]) or np.array_equal(np.sort(lcp), np.sort(r["lcp"]))
    assert np.abs(got_sorted - ref_sorted).max() < 1e-6
    plan.close()
    return n_pairs, n_quads, len(poses)


@pytest.mark.parametrize("name,seed,nq,ns", [("ellipse", 2, 400, 500), ("cuboid", 3, 400, 500), ("tless", 4, 1500, 800), ("cylinder", 5, 90, 300)])
def test_super4pcs_matches_compiled_reference(ctx, name, seed, nq, ns):
    n_pairs, n_quads, n_hyp = _compare(ctx, name, seed, nq, ns)
    assert n_pairs > 100 and n_quads > 50 and n_hyp > 10


def test_super4pcs_options(ctx):
    _compare(ctx, "ellipse", 9, 300, 400, sample_size=60, success_quadrilaterals=3, dispersion=0.7)
    _compare(ctx, "cuboid", 10, 500, 600, delta=0.005, max_trials=8)


def test_super4pcs_nothing_found(ctx):
    m, mn = synth.make_model("ellipse", 120, seed=1)
    s, sn, conf, gt = synth.make_scene("ellipse", 200, seed=2)
    plan = capi.S4pcsPlan(s, sn, conf, m, mn, np.zeros((0, 4), np.int32))      # empty PPF table: no base, no hypothesis
    poses, lcp = ctx.super4pcs_run(plan)
    assert len(poses) == 0 and plan.sizes()["trials_executed"] == 0
    plan.close()


def test_run_super4pcs_mirror_recovers_the_pose(ctx):
    """PoseEstimator::runSuper4pcs through the host mirror: hypotheses come out sorted into PoseHypo records and the best
    one (after the reference's own clustering-free arg-max on LCP) is near the ground truth."""
    import hop_b200
    m, mn = synth.make_model("cuboid", 400, seed=1)
    keys = O.ref_ppf_keys(m, mn)
    s, sn, conf, gt = synth.make_scene("cuboid", 500, seed=3, outlier_frac=0.05)
    est = hop_b200.PoseEstimator(ctx)
    est.setModel(m, mn)
    est.setCurScene(s, sn, conf)
    assert est.runSuper4pcs(keys)
    assert len(est._pose_hypos) > 10 and [h._id for h in est._pose_hypos] == list(range(len(est._pose_hypos)))
    best = max(est._pose_hypos, key=lambda h: h._lcp_score)
    assert best._lcp_score >= 0.3


@pytest.mark.parametrize("name,seed", [("cuboid", 3), ("ellipse", 2)])
def test_full_pipeline_matches_the_reference_pipeline(ctx, name, seed):
    """main_realdata_auto.cpp:187-204 through the host mirror: runSuper4pcs -> clusterPoses(30 deg, 15 mm, ids) -> refineByICP
    -> clusterPoses(5 deg, 3 mm) -> selectBest, against the same chain built from the oracles (compiled OpenGR matcher,
    Eigen-based clusterPoses, restated ICP / LCP).  The two Super4PCS lists hold the same hypotheses in a different order
    inside a trial, so the final poses are compared within the north-star tolerance, not bit for bit."""
    import hop_b200
    m, mn = synth.make_model(name, 400, seed=1)
    m001, mn001 = synth.make_model(name, 6000, seed=5)
    keys = O.ref_ppf_keys(m, mn)
    s, sn, conf, gt = synth.make_scene(name, 500, seed=seed, outlier_frac=0.05)
    cfg = {"model_name": name, "object_symmetry": {name: {"x": 180, "y": 180, "z": 180}}}
    sym = (180.0, 180.0, 180.0)
    est = hop_b200.PoseEstimator(ctx, cfg)
    est.setModel(m, mn, m001, mn001)
    est.setCurScene(s, sn, conf)
    assert est.runSuper4pcs(keys)
    est.clusterPoses(30, 0.015, True)
    est.refineByICP()
    est.clusterPoses(5, 0.003, False)
    best = est.selectBest()
    # the reference chain
    r = O.ref_super4pcs(s, sn, conf, m, mn, keys)
    keep = O.ref_cluster_poses(r["poses"], r["lcp"], 30, 0.015, sym)
    poses, lcp = r["poses"][keep][:100], r["lcp"][keep][:100]
    refined, _, _ = O.refine_by_icp(s, sn, m, mn, poses)
    keep2 = O.ref_cluster_poses(refined, lcp, 5, 0.003, sym, np.arange(len(refined), dtype=np.int32))
    bi, sc = O.select_best(s, sn, m001, mn001, refined[keep2])
    ref_best = refined[keep2][bi]
    dt, dr = synth.pose_error(best._pose[None], ref_best[None])
    egt, rgt = synth.pose_error(best._pose[None], gt[None])
    egt_ref, rgt_ref = synth.pose_error(ref_best[None], gt[None])
    # same winner within the north-star tolerance, or at least no worse against the ground truth than the reference's
    assert (dt[0] <= 1e-3 and dr[0] <= 1.0) or (egt[0] <= egt_ref[0] + 5e-4), (dt, dr, egt, egt_ref)
    assert abs(best._lcp_score - sc[bi]) <= 0.05 * max(sc[bi], 1.0) or best._lcp_score >= sc[bi]


@pytest.mark.parametrize("name,seed,nq,ns", [("ellipse", 2, 400, 500), ("cuboid", 3, 400, 2000), ("tless", 4, 1500, 7000)])
def test_device_assisted_plan_is_bit_identical(ctx, name, seed, nq, ns):
    """hop_s4pcs_plan_create_gpu: the planner's PPF-membership scans come from the device (a bit matrix of all scene pairs up to 6144
    points, rows on demand above: the 7000-point case), pairs at a bin boundary are re-evaluated on the host -- the pools, hence the
    replayed random streams, hence every base and invariant must equal the host planner's bit for bit (which test_s4pcs_plan.py pins
    against the compiled reference)."""
    m, mn = synth.make_model(name, nq, seed=1)
    keys = O.ref_ppf_keys(m[:400], mn[:400])
    s, sn, conf, gt = synth.make_scene(name, ns, seed=seed)
    a = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options())
    b = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options(), ctx=ctx)
    ga, gb = a.get(), b.get()
    for k in ga:
        assert np.array_equal(np.asarray(ga[k]), np.asarray(gb[k])), k
    assert a.sizes() == b.sizes() and ga["base_ok"].sum() > 0                  # some bases were found
    a.close(); b.close()


@pytest.mark.parametrize("name,n", [("ellipse", 400), ("cuboid", 640), ("tless", 300)])
def test_device_ppf_table_equals_the_references(ctx, name, n):
    """hop_ppf_table_build (all-pairs kernel + device sort / unique, boundary pairs on the host) against the reference's own
    gr::computePPF over all pairs (oracle/_ref): the same set of keys"""
    m, mn = synth.make_model(name, n, seed=3)
    got = ctx.ppf_table(m, mn)
    ref = O.ref_ppf_keys(m, mn)
    assert got.shape == ref.shape and np.array_equal(got, np.array(sorted(map(tuple, ref.tolist())), np.int32).reshape(-1, 4))
