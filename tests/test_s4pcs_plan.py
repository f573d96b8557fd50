"""CPU tests of the host side of Super4PCS (libhop's planner) against the reference's OWN compiled matcher (oracle/_ref):
sampling, centring, diameter, and -- through the replayed std::mt19937 / std::discrete_distribution streams -- the base
and invariants of every trial must be bit-identical."""
import numpy as np
import pytest

from hop_b200 import capi, synth
from oracle import cpu_oracle as O

needs_ref = pytest.mark.skipif(O.ref() is None or not hasattr(O.ref(), "hop_ref_s4pcs_get_trials"), reason="oracle/_ref (compiled OpenGR) not built")


@needs_ref
def test_compute_ppf_matches_reference():
    rng = np.random.default_rng(0)
    for _ in range(2000):
        p1, p2 = rng.normal(0, 0.03, 3).astype(np.float32), rng.normal(0, 0.03, 3).astype(np.float32)
        n1, n2 = rng.normal(size=3).astype(np.float32), rng.normal(size=3).astype(np.float32)
        assert np.array_equal(capi.compute_ppf(p1, n1, p2, n2), O.ref_compute_ppf(p1, n1, p2, n2))
    # ties go up, truncation first: 2.5 mm -> bin 5, 2.4999 mm -> bin 0
    z, n = np.zeros(3, np.float32), np.array([0, 0, 1], np.float32)
    assert capi.compute_ppf(z, n, np.array([0.0035, 0, 0], np.float32), n)[0] == 5
    assert capi.compute_ppf(z, n, np.array([0.0024, 0, 0], np.float32), n)[0] == 0


@needs_ref
@pytest.mark.parametrize("name,seed,nq,ns", [("ellipse", 2, 400, 500), ("cuboid", 3, 400, 500), ("tless", 4, 1500, 800), ("cylinder", 5, 90, 300),
                                             ("ellipse", 7, 400, 9000)])   # (a pool beyond 4096 points: the 4th base point is searched on all host threads)
def test_plan_matches_compiled_reference(name, seed, nq, ns):
    m, mn = synth.make_model(name, nq, seed=1)
    keys = O.ref_ppf_keys(m[:400], mn[:400])
    s, sn, conf, gt = synth.make_scene(name, ns, seed=seed)
    r = O.ref_super4pcs(s, sn, conf, m, mn, keys)
    plan = capi.S4pcsPlan(s, sn, conf, m, mn, keys)
    g = plan.get()
    assert np.array_equal(g["Qc"], r["Qc"]) and np.array_equal(g["Pc"], r["Pc"])          # sampling + shuffle + centring
    assert np.array_equal(g["centroid_P"], r["centroid_P"]) and np.array_equal(g["centroid_Q"], r["centroid_Q"])
    assert g["diameter"] == r["diameter"]
    assert np.array_equal(m[g["q_ids"]] - g["centroid_Q"], g["Qc"])
    T = len(r["base_ok"])                                                                   # the reference stops after 10 successes
    assert T >= 5 and plan.sizes()["trials"] == 30
    assert np.array_equal(g["base_ok"][:T], r["base_ok"])
    ok = r["base_ok"].astype(bool)
    assert ok.sum() >= 5
    assert np.array_equal(g["bases"][:T][ok], r["bases_all"][ok])                           # RNG replay: same 4 points, same order
    assert np.array_equal(g["inv"][:T][ok], r["inv"][ok]) and np.array_equal(g["dist"][:T][ok], r["dist"][ok])
    plan.close()


@needs_ref
def test_plan_small_model_and_options():
    """Q smaller than sample_size is used whole (matchBase.hpp:406-411); dispersion / trials / seed are honoured."""
    m, mn = synth.make_model("ellipse", 80, seed=1)
    keys = O.ref_ppf_keys(m, mn)
    s, sn, conf, gt = synth.make_scene("ellipse", 300, seed=7)
    r = O.ref_super4pcs(s, sn, conf, m, mn, keys, dispersion=0.8, success_quadrilaterals=4)
    plan = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options(dispersion=0.8, success_quadrilaterals=4, max_trials=12))
    g = plan.get()
    assert plan.sizes()["trials"] == 12 and len(g["Qc"]) == 80 and np.array_equal(g["q_ids"], np.arange(80))
    T = min(len(r["base_ok"]), 12)
    ok = r["base_ok"][:T].astype(bool)
    assert np.array_equal(g["bases"][:T][ok], r["bases_all"][:T][ok]) and np.array_equal(g["inv"][:T][ok], r["inv"][:T][ok])
    plan.close()


def test_plan_degenerate_inputs():
    m, mn = synth.make_model("ellipse", 50, seed=1)
    s, sn, conf, gt = synth.make_scene("ellipse", 100, seed=2)
    plan = capi.S4pcsPlan(s, sn, conf, m, mn, np.zeros((0, 4), np.int32))    # empty PPF table: no base can be found
    g = plan.get()
    assert not g["base_ok"].any()
    plan.close()
    plan = capi.S4pcsPlan(s[:3], sn[:3], conf[:3], m, mn, np.zeros((0, 4), np.int32))
    assert not plan.get()["base_ok"].any()
    plan.close()


@pytest.mark.parametrize("name,seed,ns", [("ellipse", 2, 959), ("cuboid", 3, 700), ("tless", 4, 1200)])
def test_plan_from_a_membership_matrix_equals_the_host_plan(name, seed, ns, monkeypatch):
    """hop_s4pcs_plan_create_gpu feeds the planner a bit matrix of PPF membership and the planner then walks bit rows instead of calling the
    host formula per pair.  HOP_PLAN_HOSTBITS builds that matrix with the host formula (no GPU needed): the plan -- sample, bases,
    invariants of all 30 trials -- must be the one the per-pair path makes (which the test above pins against the compiled reference)."""
    m, mn = synth.make_model(name, 4000, seed=1)
    sub = slice(None, None, 16)
    keys = np.unique(np.array([capi.compute_ppf(m[sub][i], mn[sub][i], m[sub][j], mn[sub][j]) for i in range(0, 250, 3) for j in range(250) if i != j], np.int32), axis=0)
    s, sn, conf, gt = synth.make_scene(name, ns, seed=seed)
    host = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options(sample_size=100)).get()
    monkeypatch.setenv("HOP_PLAN_HOSTBITS", "1")
    bits = capi.S4pcsPlan(s, sn, conf, m, mn, keys, capi.s4pcs_options(sample_size=100)).get()
    assert host.keys() == bits.keys() and int(np.sum(host["base_ok"])) >= 5
    for k in host:
        assert np.array_equal(np.asarray(host[k]), np.asarray(bits[k])), k


def test_table_free_discrete_draws_equal_the_standard_library():
    """The planner draws from std::discrete_distribution's tables without building them (k_s4pcs_plan.cu: draw_discrete).  Against the
    standard library itself (tests/support/discrete_ref.cpp builds a distribution per draw, like matchBase.hpp:120-140): the same index
    stream for random weights, zero weights, weights of very different size, one entry, no entry."""
    import ctypes as C
    import os
    import subprocess
    sup = os.path.join(os.path.dirname(os.path.abspath(__file__)), "support")
    subprocess.run(["make", "-C", sup], check=True, capture_output=True)
    ref = C.CDLL(os.path.join(sup, "libdiscrete_ref.so"))
    lib = capi.load_library()
    f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    ref.hop_ref_draw_discrete.argtypes = [f32p, C.c_int, C.c_uint32, C.c_int, i32p]
    lib.hop_debug_draw_discrete.argtypes = [f32p, C.c_int, C.c_uint32, C.c_int, i32p]
    rng = np.random.default_rng(1)
    cases = [rng.random(n).astype(np.float32) for n in (2, 3, 17, 100, 959, 4096)]
    cases += [np.where(rng.random(500) < 0.6, 0, rng.random(500)).astype(np.float32)]            # many zero weights
    cases += [(rng.random(300) * 10.0 ** rng.integers(-20, 5, 300)).astype(np.float32)]          # 25 orders of magnitude
    cases += [np.full(64, 0.5, np.float32) * np.float32(0.5) ** rng.integers(0, 12, 64)]        # the planner's own: powers of the dispersion
    cases += [np.array([3.0], np.float32), np.zeros(0, np.float32)]
    for w in cases:
        w = np.ascontiguousarray(w, np.float32)
        a, b = np.zeros(4000, np.int32), np.zeros(4000, np.int32)
        ref.hop_ref_draw_discrete(w, len(w), 7, len(a), a)
        assert lib.hop_debug_draw_discrete(w, len(w), 7, len(b), b) == 0
        assert np.array_equal(a, b), len(w)
