"""The reference's unit tests run THROUGH THE C ABI on the GPU."""
import numpy as np
import pytest

from test_oracle_kats import group_fixture

pytestmark = pytest.mark.gpu
INF = float("inf")


def single_sphere_scene(rt):
    # a bound that always passes + the unit sphere of primitive.rs:125-144
    sph = np.array([[0, 0, 0, 100.0], [0, 0, 0, 1.0]], np.float32)
    return rt.Scene.from_nodes(sph, np.array([2, 2], np.uint32), (0, -1, 0), (0, 0, -4))


def test_sphere_distance_kat(rt):
    # primitive.rs:146-155 and :157-165 (Hit::missed() start)
    s = single_sphere_scene(rt)
    d, n = s.trace_rays([[2, 0, 0], [2, 0, 0]], [[-1, 0, 0], [1, 0, 0]])
    assert d[0] == 1.0 and n[0, 0] == 1.0 and n[0, 1] == 0.0 and n[0, 2] == 0.0
    assert d[1] == INF


def test_group_intersect_kat(rt):
    # group.rs:153-170
    sph, skip, pos, dirs = group_fixture()
    s = rt.Scene.from_nodes(sph, skip, (0, -1, 0), (0, 0, -4))
    d, n = s.trace_rays(pos, dirs)
    for i in (0, 1):
        assert d[i] == 1.0 and n[i, 0] == 1.0 and n[i, 2] == 0.0
    assert d[2] == INF


def test_pyramid_counts_kat(rt):
    # group.rs:172-184
    s = rt.Scene(level=8, origin=(1.0, -1.0, 0.0))
    assert s.counts() == (5461, 21845)


def test_device_scene_matches_oracle_tree(rt, oracle):
    for level in (3, 8, 9):
        sph, skip = rt.Scene(level=level).export_nodes()
        osph, oskip = oracle.Scene(level=level).flatten()
        assert np.array_equal(skip, oskip) and np.array_equal(sph.view(np.uint32), osph.view(np.uint32))


def test_scene_light_eye(rt, oracle, gpu_scene8, oracle_scene8):
    assert np.array_equal(gpu_scene8.light().view(np.uint32), oracle_scene8.light().view(np.uint32))
    assert tuple(gpu_scene8.eye()) == (0.0, 0.0, -4.0)


def test_random_rays_match_oracle(rt, oracle, gpu_scene8, oracle_scene8):
    rng = np.random.default_rng(1234)
    n = 20000
    pos = rng.uniform(-4, 4, (n, 3)).astype(np.float32)
    tgt = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32) + np.array([0, 0.3, 0], np.float32)
    d = tgt - pos
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    gd, gn = gpu_scene8.trace_rays(pos, d)
    od, on = oracle_scene8.trace_rays(pos, d)
    assert np.array_equal(gd.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(gn.view(np.uint32), on.view(np.uint32))
    assert 0.2 < np.isfinite(gd).mean() < 1.0


def test_newton_sqrt_and_reciprocal_are_ieee_exact(rt):
    """The TILE kernel's branch-free sqrt / reciprocal must equal __fsqrt_rn / __frcp_rn bit for bit."""
    for seed in (1, 2, 3, 4):
        assert rt.selftest_math(1 << 24, seed) == (0, 0, 0, 0, 0, 0)
