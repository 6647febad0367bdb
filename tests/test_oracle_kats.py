"""The reference's own unit tests, transcribed against the CPU oracle.

vec.rs:157-170, primitive.rs:146-179, group.rs:153-184, render.rs:483-499.
"""
import math

import numpy as np

INF = float("inf")


def test_vec_normalize_lengths(oracle):
    # vec.rs:157-170
    assert oracle.vec_len((2.0, 0.0, 0.0)) == 2.0
    n = oracle.vec_normalized((2.0, 0.0, 0.0))
    assert oracle.vec_len(n) == 1.0
    assert n == (1.0, 0.0, 0.0)


def test_sphere_distance_from_ray(oracle):
    # primitive.rs:146-155: unit sphere at origin, ray from (2,0,0) towards -x
    assert oracle.sphere_distance_from_ray((0, 0, 0), 1.0, (2, 0, 0), (-1, 0, 0)) == 1.0
    assert oracle.sphere_distance_from_ray((0, 0, 0), 1.0, (2, 0, 0), (1, 0, 0)) == INF


def test_sphere_intersect(oracle):
    # primitive.rs:157-172
    d, n = oracle.sphere_intersect((0, 0, 0), 1.0, 2.0, (2, 0, 0), (-1, 0, 0))
    assert d == 1.0 and n[0] == 1.0
    d, _ = oracle.sphere_intersect((0, 0, 0), 1.0, 0.5, (2, 0, 0), (-1, 0, 0))
    assert d == 0.5, "Max Distance too short"
    d, _ = oracle.sphere_intersect((0, 0, 0), 1.0, 10.0, (2, 0, 0), (1, 0, 0))
    assert d == 10.0, "r2 is shot the wrong way"


def group_fixture():
    # group.rs:118-151: two unit spheres (origin, z=2) under a bound of radius 3
    sph = np.array([[0, 0, 0, 3.0], [0, 0, 0, 1.0], [0, 0, 2.0, 1.0]], np.float32)
    skip = np.array([3, 2, 3], np.uint32)
    rays_pos = np.array([[2, 0, 0], [2, 0, 2], [2, 0, 0]], np.float32)
    rays_dir = np.array([[-1, 0, 0], [-1, 0, 0], [1, 0, 0]], np.float32)
    return sph, skip, rays_pos, rays_dir


def test_group_intersect(oracle):
    # group.rs:153-170
    sph, skip, pos, dirs = group_fixture()
    s = oracle.Scene.from_nodes(sph, skip, (0, -1, 0), (0, 0, -4))
    dist, nrm = s.trace_rays(pos, dirs)
    for i in (0, 1):
        assert dist[i] == 1.0 and nrm[i, 0] == 1.0 and nrm[i, 2] == 0.0
    assert dist[2] == INF


def test_pyramid_counts(oracle):
    # group.rs:172-184: pyramid(8, (1,-1,0), 1.0) -> 5 children, (5461 groups, 21845 items)
    s = oracle.Scene(level=8, origin=(1.0, -1.0, 0.0))
    assert s.counts() == (5461, 21845)
    sph, skip = s.flatten()
    assert len(skip) == 5461 + 21845
    # the root has 5 children: own sphere + 4 sub-pyramids
    children, j = 0, 1
    while j < skip[0]:
        children += 1
        j = skip[j]
    assert children == 5
    for level, leaves in ((2, 5), (9, 87381), (10, 349525)):
        assert oracle.Scene(level=level).counts()[1] == leaves == (4 ** level - 1) // 3


def test_level_one_is_rejected(oracle):
    # group.rs:59-60
    import pytest
    with pytest.raises(ValueError):
        oracle.Scene(level=1)


def test_scene_default_light_and_eye(oracle_scene8):
    # render.rs:154-164
    l = oracle_scene8.light()
    ref = np.array([-1, -3, 2], np.float32)
    ref = ref * np.float32(1.0) / np.sqrt(np.float32(14.0))
    assert np.allclose(l, ref, atol=1e-7)
    assert abs(math.sqrt(float((l.astype(np.float64) ** 2).sum())) - 1.0) < 1e-6
    assert tuple(oracle_scene8.eye()) == (0.0, 0.0, -4.0)


def test_basic_rendering_shape(oracle_scene8):
    # render.rs:466-481: 64x128 spp 2 renders; every pixel is written
    img, ctr = oracle_scene8.render(64, 128, 2, threads=1)
    assert img.shape == (128, 64, 4)
    assert ctr.primary_rays == 64 * 128 * 4
    # alpha is the lit fraction: 0 on background
    assert tuple(img[0, 0]) == (34, 10, 10, 0)


def test_render_region_equals_full_frame_crop(oracle_scene8):
    full, _ = oracle_scene8.render(96, 80, 2, threads=2)
    reg, _ = oracle_scene8.render_region(96, 80, 2, 10, 20, 70, 77)
    assert np.array_equal(reg, full[20:77, 10:70])


def test_interleaved_rows_reassemble(oracle_scene8):
    full, cf = oracle_scene8.render(80, 50, 1, threads=2)
    shadow = 0
    for g in range(3):
        rows = (50 - g + 2) // 3
        band, c = oracle_scene8.render_rows(80, 50, 1, g, 3, rows, threads=2)
        assert np.array_equal(band, full[g::3])
        shadow += c.shadow_rays
    assert shadow == cf.shadow_rays


def test_identity_camera_is_the_reference_camera(oracle, oracle_scene8):
    cam = oracle.make_camera((0.0, 0.0, -4.0))
    a, _ = oracle_scene8.render(120, 90, 2, threads=2)
    b, _ = oracle_scene8.render(120, 90, 2, threads=2, camera=cam)
    assert np.array_equal(a, b)


def test_spp_zero_is_black(oracle_scene8):
    # render.rs:219-220,249-252: recip(0)=inf, 0*inf=NaN, `NaN as u8` = 0
    img, _ = oracle_scene8.render(16, 8, 0, threads=1)
    assert not img.any()
