"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every declared symbol, refuses to run
without a GPU (no CPU fallback), and the host-side helpers behave."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import hop_b200
from hop_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = hop_b200.load_library()
    names = hop_b200.declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libhop.so does not export {n}"
    out = subprocess.run(["nm", "-D", "--defined-only", hop_b200.lib_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_struct_layouts_match_header():
    assert C.sizeof(capi.PoseRec) == 80 and capi.POSE_REC_DTYPE.itemsize == 80
    assert C.sizeof(capi.IcpParams) == 40
    assert C.sizeof(capi.LcpParams) == 24
    L = hop_b200.load_library()
    p = capi.IcpParams()
    L.hop_default_icp_params(C.byref(p))
    assert (p.max_iter, p.mode, p.solver) == (10, 0, 0) and abs(p.max_dist - 0.01) < 1e-9 and p.abs_mse_eps == 1e-6
    q = capi.LcpParams()
    L.hop_default_lcp_params(C.byref(q))
    assert abs(q.dist - 0.001) < 1e-9 and q.angle_deg == 10 and (q.use_normal, q.use_dot_score, q.use_reciprocal) == (1, 1, 1)


def _have_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True).returncode == 0
    except FileNotFoundError:
        return False


@pytest.mark.skipif(_have_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(hop_b200.HopError) as e:
        hop_b200.Context(0)
    assert "no CPU path" in str(e.value) or "no CUDA device" in str(e.value)


def test_pose_layout_roundtrip():
    rng = np.random.default_rng(0)
    P = rng.normal(size=(5, 4, 4)).astype(np.float32)
    flat = capi.poses_to_colmajor(P)
    assert flat.shape == (5, 16)
    assert flat[2, 12] == P[2, 0, 3] and flat[2, 1] == P[2, 1, 0]  # column-major: translation at 12..14
    assert np.array_equal(capi.colmajor_to_poses(flat), P)


def test_synthetic_fixtures_are_deterministic_and_sized():
    for name in synth.MODELS:
        a, an = synth.make_model(name, 1234, seed=9)
        b, bn = synth.make_model(name, 1234, seed=9)
        assert a.shape == (1234, 3) and np.array_equal(a, b) and np.array_equal(an, bn)
        assert np.allclose(np.linalg.norm(an, axis=1), 1, atol=1e-5)
    s, sn, conf, gt = synth.make_scene("tless", 777, seed=4)
    assert s.shape == (777, 3) and conf.min() >= 0.8 and 0.29 < gt[2, 3] < 0.41
    hyp = synth.make_hypotheses(gt, 100, seed=1)
    assert hyp.shape == (100, 4, 4)
    R = hyp[:, :3, :3]
    assert np.allclose(np.einsum("nij,nkj->nik", R, R), np.eye(3), atol=1e-5)
    wl = synth.workload("C2")
    assert (wl["n_scene"], wl["n_model"], wl["H"]) == (2000, 10000, 1024)
