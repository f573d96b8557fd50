"""Error behaviour of the entry points added for the SURVEY 8f rows, through the raw C ABI: bad arguments come back as negative
HOP_E* codes with a message in hop_last_error -- no exception crosses the boundary, nothing is written, the context stays usable."""
import ctypes as C

import numpy as np
import pytest

from hop_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _err(ctx):
    return ctx.L.hop_last_error(ctx.h).decode()


def test_bad_arguments_return_codes(ctx):
    L, h = ctx.L, ctx.h
    V, F = synth.make_mesh("cuboid", 1)
    mesh = ctx.upload_mesh(V, F)
    cloud = ctx.upload_cloud(V, None)
    out = C.c_void_p()
    # meshes
    assert L.hop_mesh_upload(h, capi._ptr(V), 2, capi._ptr(F), len(F), C.byref(out)) < 0 and "bad arguments" in _err(ctx)
    bad = (F + 50).astype(np.int32)
    assert L.hop_mesh_upload(h, capi._ptr(V), len(V), capi._ptr(bad), len(bad), C.byref(out)) < 0 and "out of range" in _err(ctx)
    assert L.hop_sdf_query(h, None, capi._ptr(V), len(V), None, 1, None, None, None, None, None) < 0
    assert L.hop_sdf_query(h, mesh.handle, None, 5, None, 1, None, None, None, None, None) < 0
    # collision: no object mesh / no params
    p = ctx.collision_params(synth.make_collision_case("cuboid", H=2, seed=1)["params"])
    poses = capi.poses_to_colmajor(np.stack([np.eye(4)] * 2))
    keep = np.zeros(2, np.int32)
    assert L.hop_reject_by_collision(h, None, None, None, None, None, None, capi._ptr(poses), 2, C.byref(p), capi._ptr(keep), None, None) < 0
    assert L.hop_reject_by_collision(h, mesh.handle, None, None, None, None, None, capi._ptr(poses), 2, None, capi._ptr(keep), None, None) < 0
    assert L.hop_reject_by_collision(h, mesh.handle, None, None, None, None, None, capi._ptr(poses), 2, C.byref(p), None, None, None) < 0
    # render: bad camera, missing depth image, bad object mesh
    rp = ctx.render_params(width=64, height=48)
    depth = np.ones((48, 64), np.float32)
    scene_h = C.c_void_p()
    badp = ctx.render_params(width=64, height=48, z_near=0.0)
    assert L.hop_render_scene_create(h, C.byref(badp), capi._ptr(depth), None, 0, None, 0, C.byref(scene_h)) < 0 and "camera" in _err(ctx)
    assert L.hop_render_scene_create(h, C.byref(rp), None, None, 0, None, 0, C.byref(scene_h)) < 0
    assert L.hop_render_scene_create(h, C.byref(rp), capi._ptr(depth), None, 0, None, 0, C.byref(scene_h)) == 0
    wr = np.zeros(2, np.float32)
    assert L.hop_reject_by_render(h, scene_h, capi._ptr(V), len(V), capi._ptr(bad), len(bad), capi._ptr(poses), 2, capi._ptr(wr), None, None) < 0 and "out of range" in _err(ctx)
    assert L.hop_reject_by_render(h, None, capi._ptr(V), len(V), capi._ptr(F), len(F), capi._ptr(poses), 2, capi._ptr(wr), None, None) < 0
    assert L.hop_reject_by_render(h, scene_h, capi._ptr(V), len(V), capi._ptr(F), len(F), capi._ptr(poses), 2, capi._ptr(wr), None, None) == 0   # still usable
    L.hop_render_scene_destroy(h, scene_h)
    # cloud filters: the output must not be the input; parameter ranges
    same = C.c_void_p(cloud.handle.value)
    for call in (lambda: L.hop_cloud_voxel_grid(h, cloud.handle, C.c_float(0.01), C.byref(same)),
                 lambda: L.hop_cloud_pass_through(h, cloud.handle, 0, C.c_float(0), C.c_float(1), C.byref(same)),
                 lambda: L.hop_cloud_radius_outlier_removal(h, cloud.handle, C.c_float(0.01), 3, C.byref(same)),
                 lambda: L.hop_cloud_statistical_outlier_removal(h, cloud.handle, 5, C.c_float(1.0), C.byref(same))):
        assert call() < 0 and "bad arguments" in _err(ctx)
    fresh = C.c_void_p()
    assert L.hop_cloud_voxel_grid(h, cloud.handle, C.c_float(-1.0), C.byref(fresh)) < 0
    assert L.hop_cloud_pass_through(h, cloud.handle, 3, C.c_float(0), C.c_float(1), C.byref(fresh)) < 0
    assert L.hop_cloud_statistical_outlier_removal(h, cloud.handle, 65, C.c_float(1.0), C.byref(fresh)) < 0
    # hand-side entry points
    hp = ctx.hand_removal_params(np.eye(4), np.eye(4), np.eye(4), 0.0, 0.003)
    links = (C.c_void_p * 17)()
    kinds = np.zeros(17, np.int32)
    assert L.hop_remove_hand_points(h, cloud.handle, links, capi._ptr(kinds), 17, C.byref(hp), C.byref(fresh)) < 0 and "16 links" in _err(ctx)
    best = C.c_int32(0)
    assert L.hop_adjust_hand_height(h, cloud.handle, cloud.handle, None, 13, None, C.byref(best)) < 0
    # the context still works after all of that
    r = ctx.sdf_query(mesh, np.zeros((1, 3), np.float32))
    assert r["S"][0, 0] < 0
    mesh.free(); cloud.free()
