"""Parity of the CUDA path against the CPU oracle, through the C ABI.

Bar (BASELINE.json): identical hit/miss mask and >= 99.9 % of pixels within 1 LSB.
What is asserted here is stricter: every byte equal (the kernels issue the
reference's f32 operations one by one), and every per-sample classification equal.
"""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "rtrace_output_1024x768.json")))
DERIVED = json.load(open(os.path.join(HERE, "golden", "oracle_derived.json")))


def variants(rt):
    return [rt.VARIANT_LANE, rt.VARIANT_WARP, rt.VARIANT_TILE, rt.VARIANT_PHASED]


def assert_same(gpu, ref, what=""):
    if not np.array_equal(gpu, ref):
        diff = np.abs(gpu.astype(np.int16) - ref.astype(np.int16)).max(axis=-1)
        ys, xs = np.nonzero(diff)
        raise AssertionError("%s: %d of %d pixels differ (max %d), first at row %d col %d: gpu %s ref %s" % (
            what, len(ys), diff.size, diff.max(), ys[0], xs[0], gpu[ys[0], xs[0]], ref[ys[0], xs[0]]))


@pytest.mark.parametrize("w,h,spp,level", [(64, 128, 2, 8), (160, 120, 1, 8), (160, 120, 3, 8), (200, 150, 4, 5),
                                          (97, 61, 2, 9), (256, 144, 1, 10), (33, 7, 1, 8), (1, 1, 1, 8), (8, 4, 5, 2),
                                          (320, 180, 5, 7), (200, 120, 6, 8), (131, 77, 7, 6), (256, 144, 8, 8),
                                          (40, 30, 9, 5),    # 5 .. 8: PHASED's widest templates; 9: per-lane walk only
                                          (96, 54, 1, 11), (48, 27, 2, 12)])   # levels 11 / 12: per-lane walk only
def test_frame_matches_oracle(rt, oracle, w, h, spp, level):
    gs, os_ = rt.Scene(level=level), oracle.Scene(level=level)
    ref, ctr = os_.render(w, h, spp)
    for v in variants(rt):
        rt.set_variant(v)
        img = rt.Renderer.render(rt.RenderOptions(w, h, spp), gs)
        assert_same(img, ref, "variant %d %dx%d spp %d L%d" % (v, w, h, spp, level))
        assert gs.count_rays(w, h, spp) == (ctr.primary_rays, ctr.shadow_rays)
    rt.set_variant(rt.VARIANT_AUTO)


def test_per_sample_classification_matches_oracle(rt, oracle, gpu_scene8, oracle_scene8):
    w, h, spp = 192, 144, 2
    ref, okinds, _ = oracle_scene8.render_region(w, h, spp, 0, 0, w, h, kinds=True)
    for v in variants(rt):
        rt.set_variant(v)
        img, kinds = rt.Renderer.render_rows(rt.RenderOptions(w, h, spp), gpu_scene8, kinds=True)
        assert np.array_equal(kinds, okinds), "hit/miss/shadow mask differs for variant %d" % v
        assert_same(img, ref)
    rt.set_variant(rt.VARIANT_AUTO)
    assert set(np.unique(okinds)) == {0, 1, 2, 3}


def test_make_image_matches_reference_golden(rt, gpu_scene8):
    """C1 (`make image`): the GPU frame equals the reference's shipped PNG, row by row."""
    w, h, spp = GOLD["width"], GOLD["height"], GOLD["spp"]
    for v in variants(rt):
        rt.set_variant(v)
        img = rt.Renderer.render(rt.RenderOptions(w, h, spp), gpu_scene8)
        rgb = np.ascontiguousarray(img[:, :, :3])
        bad = [y for y in range(h) if zlib.crc32(rgb[y].tobytes()) != GOLD["row_crc32"][y]]
        assert not bad, "variant %d: rows differing from the reference image: %s" % (v, bad[:10])
        assert hashlib.sha256(rgb.tobytes()).hexdigest() == GOLD["rgb_sha256"]
    rt.set_variant(rt.VARIANT_AUTO)
    assert gpu_scene8.count_rays(w, h, spp) == (12582912, 7211901)


BASELINE_CASES = [c for c in DERIVED["cases"] if c["width"] * c["height"] >= 1024 * 768]


@pytest.mark.parametrize("case", BASELINE_CASES,
                         ids=lambda c: "%dx%d_spp%d_L%d" % (c["width"], c["height"], c["spp"], c["level"]))
def test_baseline_frames_match_oracle_fixture(rt, case):
    """Every BASELINE configuration, FULL frame, against the committed oracle hash: C1 (1024x768, 4x4, level 8),
    C2 (3840x2160, 1 spp) at levels 8/9/10, C3 (3840x2160, 4x4, level 9 = C5's frame 0) and C4 (7680x4320, 4x4,
    level 9) -- plus the ray counts, counted on the device the way the reference's work is counted."""
    gs = rt.Scene(level=case["level"])
    img, st = rt.Renderer.render(rt.RenderOptions(case["width"], case["height"], case["spp"]), gs, want_stats=True)
    assert hashlib.sha256(img.tobytes()).hexdigest() == case["rgba_sha256"]
    assert st.variant_used in (rt.VARIANT_TILE, rt.VARIANT_PHASED)   # a candidate-list variant ran (no silent LANE fallback)
    p, s = gs.count_rays(case["width"], case["height"], case["spp"])
    assert (p, s) == (case["counters"]["primary_rays"], case["counters"]["shadow_rays"])


def test_c5_real_size_orbit_frames_match_oracle_fixture(rt):
    """BASELINE C5 at real size: orbit frames 7, 41, 88 of the 120-frame sweep (3840x2160, 4x4, level 9) through
    rt_render_sweep / rt_render_sweep_rgb, against the committed oracle hashes."""
    cases = DERIVED["c5_cases"]
    w, h, spp, level = cases[0]["width"], cases[0]["height"], cases[0]["spp"], cases[0]["level"]
    gs = rt.Scene(level=level)
    cams = [rt.orbit_camera(c["frame"], c["n_frames"]) for c in cases]
    got, rgb = {}, {}
    rt.Renderer.render_sweep(rt.RenderOptions(w, h, spp), gs, len(cams), cameras=cams,
                             on_frame=lambda i, a: got.__setitem__(i, hashlib.sha256(a.tobytes()).hexdigest()))
    rt.Renderer.render_sweep(rt.RenderOptions(w, h, spp), gs, len(cams), cameras=cams, rgb=True,
                             on_frame=lambda i, a: rgb.__setitem__(i, hashlib.sha256(a.tobytes()).hexdigest()))
    for i, c in enumerate(cases):
        assert got[i] == c["rgba_sha256"], "orbit frame %d" % c["frame"]
        assert rgb[i] == c["rgb_sha256"], "orbit frame %d (RGB8)" % c["frame"]
        assert gs.count_rays(w, h, spp, camera=cams[i]) == (c["counters"]["primary_rays"], c["counters"]["shadow_rays"])


def test_render_region_semantics(rt, oracle, gpu_scene8, oracle_scene8):
    """Renderer::render_region (render.rs:218-255): any window of the image, row-major from row b."""
    o = rt.RenderOptions(96, 80, 2)
    full, _ = oracle_scene8.render(96, 80, 2)
    for (l, b, r, t) in [(0, 0, 96, 80), (10, 20, 70, 77), (0, 64, 64, 80), (95, 79, 96, 80), (5, 5, 5, 9)]:
        reg = rt.Renderer.render_region(o, gpu_scene8, l, b, r, t)
        assert reg.shape == (t - b, r - l, 4)
        assert np.array_equal(reg, full[b:t, l:r])
    with pytest.raises(rt.RtError) as e:
        rt.Renderer.render_region(o, gpu_scene8, 0, 0, 97, 80)
    assert e.value.code == rt.RT_ERR_INVALID


def test_buckets_tile_the_frame(rt, oracle_scene8, gpu_scene8):
    """The reference's schedule (render.rs:273-298): 64x64 buckets, each a render_region call, reassemble
    the `make image`-shaped frame; a bucket's cost is its own area (the per-lane kernel takes the window)."""
    w, h, spp = 256, 192, 4
    o = rt.RenderOptions(w, h, spp)
    full, _ = oracle_scene8.render(w, h, spp)
    got = np.zeros_like(full)
    for y in range(0, h, 64):
        for x in range(0, w, 64):
            got[y:y + 64, x:x + 64] = rt.Renderer.render_region(o, gpu_scene8, x, y, x + 64, y + 64)
    assert_same(got, full, "64x64 buckets")


@pytest.mark.parametrize("w,h,step", [(97, 61, 1), (97, 61, 4), (200, 150, 7), (64, 64, 64), (50, 40, 100)])
def test_undersampled_preview(rt, oracle_scene8, gpu_scene8, w, h, step):
    """rt_render_preview: every step x step block carries the reference's 1-spp value of its first pixel."""
    full, _ = oracle_scene8.render(w, h, 1)
    ys, xs = (np.arange(h) // step) * step, (np.arange(w) // step) * step
    got = rt.Renderer.render_preview(rt.RenderOptions(w, h, 1), gpu_scene8, step)
    assert_same(got, full[ys][:, xs], "preview step %d" % step)
    with pytest.raises(rt.RtError) as e:
        rt.Renderer.render_preview(rt.RenderOptions(w, h, 1), gpu_scene8, 0)
    assert e.value.code == rt.RT_ERR_INVALID


def test_interleaved_rows_and_pitch(rt, oracle_scene8, gpu_scene8):
    """The multi-GPU partition: rows g, g+G, ... rendered independently reassemble the frame."""
    w, h, spp, G = 120, 67, 2, 4
    full, _ = oracle_scene8.render(w, h, spp)
    frame = np.zeros((h, w, 4), np.uint8)
    for g in range(G):
        band = rt.Renderer.render_rows(rt.RenderOptions(w, h, spp), gpu_scene8, row_start=g, row_stride=G)
        assert np.array_equal(band, full[g::G])
        frame[g::G] = band
    assert np.array_equal(frame, full)
    # host pitch wider than the row
    padded = np.zeros((h, w + 8, 4), np.uint8)
    rt.Renderer.render_rows(rt.RenderOptions(w, h, spp), gpu_scene8, out_ptr=padded.ctypes.data, pitch=(w + 8) * 4,
                            row_count=h)
    assert np.array_equal(padded[:, :w], full) and not padded[:, w:].any()


def test_orbit_camera_matches_oracle(rt, oracle, gpu_scene8, oracle_scene8):
    w, h, spp = 160, 90, 2
    base, _ = oracle_scene8.render(w, h, spp)
    for f in (0, 7, 30, 61):
        gc, oc = rt.orbit_camera(f, 120), oracle.Camera()
        for k in ("eye", "right", "up", "forward"):
            getattr(oc, k)[:] = getattr(gc, k)[:]
        ref, _ = oracle_scene8.render(w, h, spp, camera=oc)
        img = rt.Renderer.render(rt.RenderOptions(w, h, spp), gpu_scene8, camera=gc)
        assert_same(img, ref, "orbit frame %d" % f)
        if f == 0:
            assert np.array_equal(img, base)  # frame 0 is the reference camera


def test_device_output_and_stats(rt, oracle_scene8, gpu_scene8):
    torch = pytest.importorskip("torch")
    w, h, spp = 256, 100, 1
    ref, ctr = oracle_scene8.render(w, h, spp)
    fb = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    _, st = rt.Renderer.render_rows(rt.RenderOptions(w, h, spp), gpu_scene8, out_ptr=fb.data_ptr(),
                                    stream=torch.cuda.current_stream().cuda_stream, want_stats=True)
    torch.cuda.synchronize()
    assert np.array_equal(fb.cpu().numpy(), ref)
    assert st.primary_rays == ctr.primary_rays and st.kernel_ms > 0 and st.kernel_launches in (1, 4)


def test_spp_zero_and_bad_arguments(rt, gpu_scene8):
    img = rt.Renderer.render(rt.RenderOptions(16, 8, 0), gpu_scene8)
    assert not img.any()  # render.rs:219-220: 0 * inf = NaN -> 0
    for bad in (rt.RenderOptions(0, 8, 1), rt.RenderOptions(8, 0, 1), rt.RenderOptions(70000, 8, 1)):
        with pytest.raises(rt.RtError) as e:
            rt.Renderer.render_rows(bad, gpu_scene8, row_count=1)
        assert e.value.code == rt.RT_ERR_INVALID
    with pytest.raises(rt.RtError):
        rt.Renderer.render_rows(rt.RenderOptions(8, 8, 1), gpu_scene8, row_start=4, row_stride=2, row_count=3)
    with pytest.raises(rt.RtError) as e:
        rt.Scene(level=1)
    assert e.value.code == rt.RT_ERR_INVALID


def test_full_size_properties_c3(rt, oracle):
    """C3-sized frame (3840x2160, spp 4, level 9): size-independent properties + sampled rows vs the oracle."""
    w, h, spp, level = 3840, 2160, 4, 9
    gs, os_ = rt.Scene(level=level), oracle.Scene(level=level)
    o = rt.RenderOptions(w, h, spp)
    full = rt.Renderer.render(o, gs)
    # idempotence
    assert np.array_equal(full, rt.Renderer.render(o, gs))
    # the interleaved partition reassembles the frame
    for g in (0, 3):
        assert np.array_equal(rt.Renderer.render_rows(o, gs, row_start=g, row_stride=8), full[g::8])
    # the two kernel variants agree on every byte
    rt.set_variant(rt.VARIANT_LANE)
    lane = rt.Renderer.render_rows(o, gs, row_start=1000, row_stride=1, row_count=64)
    rt.set_variant(rt.VARIANT_AUTO)
    assert np.array_equal(lane, full[1000:1064])
    # sampled rows against the oracle
    rows = [0, 411, 1080, 1333, 1600, 2159]
    for y in rows:
        ref, _ = os_.render_rows(w, h, spp, y, 1, 1)
        assert_same(full[y:y + 1], ref, "row %d" % y)


def test_sweep_delivers_every_frame_in_order(rt, oracle, gpu_scene8, oracle_scene8):
    """rt_render_sweep: double-buffered frames, each equal to the oracle's frame for that camera."""
    w, h, spp, n = 160, 90, 1, 5
    cams = [rt.orbit_camera(f, 40) for f in range(n)]
    got = {}
    st = rt.Renderer.render_sweep(rt.RenderOptions(w, h, spp), gpu_scene8, n, cameras=cams,
                                  on_frame=lambda f, a: got.__setitem__(f, a.copy()))
    assert sorted(got) == list(range(n)) and st.primary_rays == w * h * n
    for f in range(n):
        oc = oracle.Camera()
        for k in ("eye", "right", "up", "forward"):
            getattr(oc, k)[:] = getattr(cams[f], k)[:]
        ref, _ = oracle_scene8.render(w, h, spp, camera=oc)
        assert_same(got[f], ref, "sweep frame %d" % f)
    # RGB sweep: the same frames without the alpha byte (odd pixel count exercises the tail)
    rgbs = {}
    rt.Renderer.render_sweep(rt.RenderOptions(w, h, spp), gpu_scene8, n, cameras=cams, rgb=True,
                             on_frame=lambda f, a: rgbs.__setitem__(f, a.copy()))
    assert all(np.array_equal(rgbs[f], got[f][:, :, :3]) for f in range(n))
    odd = {}
    rt.Renderer.render_sweep(rt.RenderOptions(33, 7, 1), gpu_scene8, 1, rgb=True, on_frame=lambda f, a: odd.__setitem__(f, a.copy()))
    assert np.array_equal(odd[0], oracle_scene8.render(33, 7, 1)[0][:, :, :3])
    # reference camera when no cameras are given
    got.clear()
    rt.Renderer.render_sweep(rt.RenderOptions(w, h, spp), gpu_scene8, 3, on_frame=lambda f, a: got.__setitem__(f, a.copy()))
    base, _ = oracle_scene8.render(w, h, spp)
    assert all(np.array_equal(got[f], base) for f in range(3))


def test_auto_picks_a_variant_for_every_regime(rt, oracle):
    """AUTO's three regimes (noise-limited leaves -> LANE, small frame -> TILE, large -> PHASED) all match the oracle."""
    for (w, h, spp, level) in [(640, 360, 1, 10), (640, 360, 2, 8), (2560, 1440, 1, 8)]:
        gs, os_ = rt.Scene(level=level), oracle.Scene(level=level)
        rt.set_variant(rt.VARIANT_AUTO)
        img = rt.Renderer.render(rt.RenderOptions(w, h, spp), gs)
        ref, _ = os_.render(w, h, spp)
        assert_same(img, ref, "auto %dx%d spp %d L%d" % (w, h, spp, level))


def test_candidate_pool_overflow_falls_back_to_the_lane_walk(rt, oracle_scene8, gpu_scene8):
    """A cull tile whose candidates do not fit the pool is rendered by the per-lane walk: same bytes."""
    w, h, spp = 512, 288, 1
    ref, _ = oracle_scene8.render(w, h, spp)
    rt.set_variant(rt.VARIANT_PHASED)
    try:
        for units in ("64", "600", "3000"):   # nothing fits / some tiles fit / most tiles fit
            os.environ["RTRACE_POOL_UNITS"] = units
            img, kinds = rt.Renderer.render_rows(rt.RenderOptions(w, h, spp), gpu_scene8, kinds=True)
            assert_same(img, ref, "pool of %s units" % units)
    finally:
        os.environ.pop("RTRACE_POOL_UNITS", None)
        rt.set_variant(rt.VARIANT_AUTO)
    img = rt.Renderer.render(rt.RenderOptions(w, h, spp), gpu_scene8)   # pool back to its normal size
    assert_same(img, ref)


def test_c4_sized_frame_agrees_with_the_exact_lane_walk(rt):
    """BASELINE C4 (7680x4320, 4x4, level 9): bands of the candidate-list frame equal the per-lane walk,
    which is the reference recursion restated (and is itself checked against the oracle above)."""
    w, h, spp = 7680, 4320, 4
    gs = rt.Scene(level=9)
    o = rt.RenderOptions(w, h, spp)
    rt.set_variant(rt.VARIANT_AUTO)
    full = rt.Renderer.render(o, gs)
    rt.set_variant(rt.VARIANT_LANE)
    try:
        for y0 in (1200, 2177, 3900):
            band = rt.Renderer.render_rows(o, gs, row_start=y0, row_stride=1, row_count=24)
            assert np.array_equal(band, full[y0:y0 + 24]), "rows %d.." % y0
    finally:
        rt.set_variant(rt.VARIANT_AUTO)
    # interleaved partition of the big frame (what 8 GPUs would each render)
    part = rt.Renderer.render_rows(o, gs, row_start=5, row_stride=8)
    assert np.array_equal(part, full[5::8])


@pytest.mark.parametrize("name,kw", [
    ("light_along_view", dict(light=(0.0, 0.0, -1.0))),          # shadow rays parallel to the view axis: degenerate strip
    ("light_from_below", dict(light=(0.3, 2.0, 0.5))),
    ("light_sideways", dict(light=(-4.0, -0.2, 0.1))),
    ("eye_close", dict(eye=(0.4, 0.2, -2.2))),
    ("eye_inside_root_bound", dict(eye=(0.1, -0.2, -1.6))),
    ("eye_far_off_axis", dict(eye=(3.0, 1.5, -9.0))),
    ("scaled_shifted_scene", dict(origin=(0.7, -0.4, 0.3), radius=0.6)),
    ("big_scene", dict(origin=(0.0, -2.0, 4.0), radius=2.5, eye=(0.0, 0.0, -6.0))),
])
def test_other_scenes_match_oracle(rt, oracle, name, kw):
    """Scene::default is one point in the parameter space; the culling geometry must stay conservative
    for any light, eye and scene placement (every variant, byte for byte, plus the per-sample mask)."""
    level, w, h, spp = 6, 224, 126, 2
    gs, os_ = rt.Scene(level=level, **kw), oracle.Scene(level=level, **kw)
    ref, okinds, ctr = os_.render_region(w, h, spp, 0, 0, w, h, kinds=True)
    try:
        for v in variants(rt):
            rt.set_variant(v)
            img, kinds = rt.Renderer.render_rows(rt.RenderOptions(w, h, spp), gs, kinds=True)
            assert np.array_equal(kinds, okinds), "%s: mask differs for variant %d" % (name, v)
            assert_same(img, ref, "%s variant %d" % (name, v))
    finally:
        rt.set_variant(rt.VARIANT_AUTO)
    assert ctr.primary_hits > 0


def test_other_scenes_at_high_resolution(rt, oracle):
    """The same at a resolution where the tiles are small against the spheres (PHASED regime)."""
    kw = dict(light=(0.0, 0.0, -1.0), eye=(0.5, 0.3, -3.0))
    gs, os_ = rt.Scene(level=7, **kw), oracle.Scene(level=7, **kw)
    w, h, spp = 1920, 1080, 1
    ref, _ = os_.render(w, h, spp)
    try:
        for v in (rt.VARIANT_TILE, rt.VARIANT_PHASED):
            rt.set_variant(v)
            assert_same(rt.Renderer.render(rt.RenderOptions(w, h, spp), gs), ref, "variant %d" % v)
    finally:
        rt.set_variant(rt.VARIANT_AUTO)


def test_blocked_rows_and_absolute_addressing(rt, oracle_scene8, gpu_scene8):
    """rt_render_row_blocks: rank r of N renders blocks of B rows, N blocks apart, and stores each row
    at its image row of a whole frame -- the ranks' stores together are the gathered frame."""
    torch = pytest.importorskip("torch")
    from rtrace_b200 import partition
    w, h, spp = 200, 150, 2   # 150 is not a multiple of the block: the last block is partial
    ref, _ = oracle_scene8.render(w, h, spp)
    o = rt.RenderOptions(w, h, spp)
    for world, block in ((3, 16), (2, 8), (8, 16)):
        frame = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        seen = []
        for rank in range(world):
            start, stride, blk, count = partition.block_band_spec(h, rank, world, block)
            seen += partition.block_band_rows(h, rank, world, block)
            if count:
                rt.Renderer.render_row_blocks(o, gpu_scene8, start, stride, blk, count, frame.data_ptr(), pitch=w * 4,
                                              absolute_rows=True, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert sorted(seen) == list(range(h))
        assert np.array_equal(frame.cpu().numpy(), ref), "world %d block %d" % (world, block)
    # dense band (absolute_rows off): local row j is block row j
    start, stride, blk, count = partition.block_band_spec(h, 1, 3, 16)
    band = torch.zeros((count, w, 4), dtype=torch.uint8, device="cuda")
    rt.Renderer.render_row_blocks(o, gpu_scene8, start, stride, blk, count, band.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(band.cpu().numpy(), ref[partition.block_band_rows(h, 1, 3, 16)])
    with pytest.raises(rt.RtError):
        rt.Renderer.render_row_blocks(o, gpu_scene8, 0, 48, 12, 10, band.data_ptr())   # block not a power of two


@pytest.mark.parametrize("kw,frames", [
    (dict(), (7, 41, 88)),                               # Scene::default from three orbit positions
    (dict(light=(-4.0, -0.2, 0.1)), (0, 23)),            # light nearly sideways
    (dict(light=(0.3, 2.0, 0.5)), (0, 100)),             # light from below
    (dict(light=(0.0, 0.0, -1.0)), (0, 60)),             # light along / against the view axis
])
def test_phased_prefilters_hold_for_cameras_and_lights(rt, oracle, kw, frames):
    """PHASED regime (tiles small against the spheres): the image-space boxes of the primary candidates
    are computed in camera coordinates and the shadow discs in the plane perpendicular to the light, so
    rotated cameras and other lights must give the oracle's bytes too."""
    w, h, level = 1280, 720, 8
    gs, os_ = rt.Scene(level=level, **kw), oracle.Scene(level=level, **kw)
    rt.set_variant(rt.VARIANT_PHASED)
    try:
        for i, f in enumerate(frames):
            spp = 1 + (i % 2)
            gc, oc = rt.orbit_camera(f, 120), oracle.Camera()
            for k in ("eye", "right", "up", "forward"):
                getattr(oc, k)[:] = getattr(gc, k)[:]
            ref, _ = os_.render(w, h, spp, camera=oc)
            img = rt.Renderer.render(rt.RenderOptions(w, h, spp), gs, camera=gc)
            assert_same(img, ref, "%s orbit frame %d spp %d" % (kw, f, spp))
    finally:
        rt.set_variant(rt.VARIANT_AUTO)


@pytest.mark.parametrize("kw", [
    dict(eye=(0.3, 0.6, -3.4)),                                          # just outside the root bound: many tiles fully covered
    dict(origin=(0.0, -2.0, 4.0), radius=2.5, eye=(0.0, 0.0, -6.0)),     # a scene 2.5x as large
    dict(eye=(3.0, 1.5, -9.0)),                                          # far and off axis: small projected leaves
    dict(light=(1.0, -0.4, 3.0)),                                        # light from behind the camera: almost everything lit
])
def test_phased_occlusion_culling_holds_for_other_eyes(rt, oracle, kw):
    """The occlusion pruning of the primary cull and the single-occluder shortcut of the shadow cull
    assume nothing about Scene::default: other eyes, scene sizes and lights give the oracle's bytes.
    (Every eye here lies outside the root bound, so the candidate-list kernels really run.)"""
    w, h, spp, level = 1600, 900, 1, 8
    gs, os_ = rt.Scene(level=level, **kw), oracle.Scene(level=level, **kw)
    ref, _ = os_.render(w, h, spp)
    rt.set_variant(rt.VARIANT_PHASED)
    try:
        assert_same(rt.Renderer.render(rt.RenderOptions(w, h, spp), gs), ref, str(kw))
        tiles = rt.debug_phased_tiles(gs)
        assert tiles.shape[1] == 2 and int(tiles[:, 0][tiles[:, 0] != 0xffffffff].sum()) > 0
    finally:
        rt.set_variant(rt.VARIANT_AUTO)


@pytest.mark.parametrize("kw,expect_lane", [
    (dict(origin=(1000.0, -1.0, 0.0), eye=(1000.0, 0.0, -4.0)), True),     # far from the coordinate origin
    (dict(origin=(-300.0, 200.0, 150.0), radius=3.0, eye=(-300.0, 203.0, 130.0)), True),
    (dict(origin=(9.0, -1.0, 3.0), eye=(9.0, 0.0, -1.0)), False),          # translated, still inside the analysed range
])
def test_translated_scenes_and_the_variant_that_ran(rt, oracle, kw, expect_lane):
    """The cull pre-filters carry absolute slacks sized for coordinates below 16 (rt_api.cpp,
    within_analysed_range): a scene placed far from the origin must take the per-lane walk -- and says so in
    rt_stats.variant_used -- instead of trusting slacks its coordinate rounding has outgrown; a moderately
    translated scene stays on the candidate-list path.  Both give the oracle's bytes."""
    level, w, h, spp = 7, 1280, 720, 1
    gs, os_ = rt.Scene(level=level, **kw), oracle.Scene(level=level, **kw)
    ref, _ = os_.render(w, h, spp)
    for v in (rt.VARIANT_AUTO, rt.VARIANT_PHASED, rt.VARIANT_TILE):
        rt.set_variant(v)
        try:
            img, st = rt.Renderer.render(rt.RenderOptions(w, h, spp), gs, want_stats=True)
        finally:
            rt.set_variant(rt.VARIANT_AUTO)
        assert_same(img, ref, "%s variant %d" % (kw, v))
        assert (st.variant_used == rt.VARIANT_LANE) == expect_lane, (kw, v, st.variant_used)


def test_variant_used_reports_the_fallbacks(rt, gpu_scene8):
    """rt_stats.variant_used: AUTO resolved, and LANE whenever the arguments leave the candidate-list domain
    (samples per pixel beyond the templates, an eye inside the root bound)."""
    def used(scene, w, h, spp, variant=None, camera=None):
        rt.set_variant(rt.VARIANT_AUTO if variant is None else variant)
        try:
            return rt.Renderer.render(rt.RenderOptions(w, h, spp), scene, camera=camera, want_stats=True)[1].variant_used
        finally:
            rt.set_variant(rt.VARIANT_AUTO)
    assert used(gpu_scene8, 2560, 1440, 1) == rt.VARIANT_PHASED
    assert used(gpu_scene8, 640, 360, 2) in (rt.VARIANT_TILE, rt.VARIANT_PHASED)
    assert used(gpu_scene8, 64, 64, 9, rt.VARIANT_PHASED) == rt.VARIANT_LANE          # spp beyond the fast path
    assert used(gpu_scene8, 1280, 720, 5) == rt.VARIANT_PHASED and used(gpu_scene8, 960, 540, 8) == rt.VARIANT_PHASED
    assert used(gpu_scene8, 64, 64, 6, rt.VARIANT_TILE) == rt.VARIANT_LANE            # the fused TILE kernel stops at 4x4
    assert used(gpu_scene8, 64, 64, 1, rt.VARIANT_WARP) == rt.VARIANT_WARP
    inside = rt.Scene(level=6, eye=(0.1, -0.2, -1.6))
    assert used(inside, 1280, 720, 1, rt.VARIANT_PHASED) == rt.VARIANT_LANE


def test_sweep_pull_draws_frames_from_a_queue(rt, oracle, gpu_scene8, oracle_scene8):
    """rt_render_sweep_pull: the library asks for the next frame whenever its pipeline has room; ids chosen by
    the caller come back with the frames (in pull order); a None camera is the reference camera; the shared
    counter helper is a fetch-and-add."""
    w, h, spp = 160, 90, 2
    cams = {f: rt.orbit_camera(f, 40) for f in (3, 11, 27)}
    plan = [(103, cams[3]), (7, None), (111, cams[11]), (127, cams[27]), (9, None)]
    it = iter(plan)
    got = []
    st = rt.Renderer.render_sweep_pull(rt.RenderOptions(w, h, spp), gpu_scene8, lambda: next(it, None),
                                       on_frame=lambda f, a: got.append((f, a.copy())))
    assert [f for f, _ in got] == [p[0] for p in plan] and st.primary_rays == w * h * spp * spp * len(plan)
    base, _ = oracle_scene8.render(w, h, spp)
    for (fid, img), (_, cam) in zip(got, plan):
        if cam is None:
            assert np.array_equal(img, base)
        else:
            oc = oracle.Camera()
            for k in ("eye", "right", "up", "forward"):
                getattr(oc, k)[:] = getattr(cam, k)[:]
            assert_same(img, oracle_scene8.render(w, h, spp, camera=oc)[0], "pulled frame %d" % fid)
    # an empty queue renders nothing
    st = rt.Renderer.render_sweep_pull(rt.RenderOptions(w, h, spp), gpu_scene8, lambda: None, rgb=True)
    assert st.primary_rays == 0
    import ctypes
    word = ctypes.c_uint64(5)
    assert rt.atomic_fetch_add_u64(ctypes.addressof(word), 3) == 5 and word.value == 8
