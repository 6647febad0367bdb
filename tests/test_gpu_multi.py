"""In-process multi-GPU entry points: rt_render_frame_multi (one frame as interleaved row blocks whose
kernels store into GPU 0's frame over NVLink) and rt_render_sweep_multi (a sweep sharded by frame).
Skipped on a 1-GPU box; `gpurun --gpus 2 -- python -m pytest tests -m gpu` runs them (logs in profiles/)."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
DERIVED = json.load(open(os.path.join(HERE, "golden", "oracle_derived.json")))


@pytest.fixture(scope="module")
def replicas(rt):
    n = rt.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")

    def make(level=8, count=None):
        scenes = []
        for g in range(min(n, count or 8)):
            rt.set_device(g)
            scenes.append(rt.Scene(level=level))
        rt.set_device(0)
        return scenes
    return make


def test_multi_gpu_frame_matches_oracle(rt, oracle_scene8, replicas):
    scenes = replicas(8, 4)
    for (w, h, spp) in [(256, 131, 2), (640, 360, 1), (33, 7, 1)]:
        img, st = rt.Renderer.render_multi(rt.RenderOptions(w, h, spp), scenes, want_stats=True)
        ref, _ = oracle_scene8.render(w, h, spp)
        assert np.array_equal(img, ref), (w, h, spp)
        assert st.gpus == len(scenes)


def test_multi_gpu_every_gpu_count(rt, oracle_scene8, replicas):
    """2 .. all GPUs of the box (odd counts too): rows of a 150-row frame never divide evenly."""
    scenes = replicas(8)
    w, h, spp = 200, 150, 2
    ref, _ = oracle_scene8.render(w, h, spp)
    for n in range(2, len(scenes) + 1):
        img = rt.Renderer.render_multi(rt.RenderOptions(w, h, spp), scenes[:n])
        assert np.array_equal(img, ref), "n = %d" % n


@pytest.mark.parametrize("w,h,spp,level", [(3840, 2160, 4, 9), (7680, 4320, 4, 9)], ids=["c3", "c4_8k"])
def test_multi_gpu_full_size_frames_match_oracle_hash(rt, replicas, w, h, spp, level):
    """BASELINE C4 (8K, 4x4, level 9) and a C3-sized frame split over all GPUs of the box: the gathered frame's
    sha256 equals the oracle's committed hash."""
    case = [c for c in DERIVED["cases"] if (c["width"], c["height"], c["spp"], c["level"]) == (w, h, spp, level)][0]
    scenes = replicas(level)
    img, st = rt.Renderer.render_multi(rt.RenderOptions(w, h, spp), scenes, want_stats=True)
    assert hashlib.sha256(img.tobytes()).hexdigest() == case["rgba_sha256"]
    assert st.gpus == len(scenes) and st.variant_used == rt.VARIANT_PHASED


def test_same_scene_twice_is_rejected(rt, gpu_scene8):
    with pytest.raises(rt.RtError) as e:
        rt.Renderer.render_multi(rt.RenderOptions(64, 64, 1), [gpu_scene8, gpu_scene8])
    assert e.value.code == rt.RT_ERR_INVALID
    with pytest.raises(rt.RtError) as e:
        rt.Renderer.render_sweep_multi(rt.RenderOptions(64, 64, 1), [gpu_scene8, gpu_scene8], 3)
    assert e.value.code == rt.RT_ERR_INVALID


def test_sweep_multi_delivers_every_frame_in_order(rt, oracle, oracle_scene8, replicas):
    """rt_render_sweep_multi: frame f on GPU f mod N, callback in frame order, every frame the oracle's."""
    scenes = replicas(8)
    w, h, spp, n = 160, 90, 2, 2 * len(scenes) + 3     # not a multiple of the GPU count
    cams = [rt.orbit_camera(f, 40) for f in range(n)]
    order, got = [], {}
    st = rt.Renderer.render_sweep_multi(rt.RenderOptions(w, h, spp), scenes, n, cameras=cams,
                                        on_frame=lambda f, a: (order.append(f), got.__setitem__(f, a.copy())))
    assert order == list(range(n)) and st.gpus == len(scenes) and st.primary_rays == w * h * spp * spp * n
    for f in range(n):
        oc = oracle.Camera()
        for k in ("eye", "right", "up", "forward"):
            getattr(oc, k)[:] = getattr(cams[f], k)[:]
        ref, _ = oracle_scene8.render(w, h, spp, camera=oc)
        assert np.array_equal(got[f], ref), "frame %d" % f
    # RGB8 delivery and fewer frames than GPUs
    rgbs = {}
    rt.Renderer.render_sweep_multi(rt.RenderOptions(w, h, spp), scenes, 1, cameras=cams[:1], rgb=True,
                                   on_frame=lambda f, a: rgbs.__setitem__(f, a.copy()))
    assert list(rgbs) == [0] and np.array_equal(rgbs[0], got[0][:, :, :3])


def test_sweep_multi_real_size_c5_frames(rt, replicas):
    """BASELINE C5 frames at real size (3840x2160, 4x4, level 9) through the frame-sharded sweep, against the
    committed oracle hashes of orbit frames 7, 41 and 88."""
    cases = DERIVED.get("c5_cases") or []
    if not cases:
        pytest.skip("no c5 fixtures")
    scenes = replicas(cases[0]["level"])
    w, h, spp = cases[0]["width"], cases[0]["height"], cases[0]["spp"]
    frames = [c["frame"] for c in cases]
    cams = [rt.orbit_camera(f, cases[0]["n_frames"]) for f in frames]
    got = {}
    rt.Renderer.render_sweep_multi(rt.RenderOptions(w, h, spp), scenes, len(cams), cameras=cams,
                                   on_frame=lambda i, a: got.__setitem__(i, hashlib.sha256(a.tobytes()).hexdigest()))
    for i, c in enumerate(cases):
        assert got[i] == c["rgba_sha256"], "orbit frame %d" % c["frame"]


def test_single_scene_multi_call_is_a_plain_frame(rt, gpu_scene8, oracle_scene8):
    img = rt.Renderer.render_multi(rt.RenderOptions(128, 96, 2), [gpu_scene8])
    ref, _ = oracle_scene8.render(128, 96, 2)
    assert np.array_equal(img, ref)
    got = {}
    rt.Renderer.render_sweep_multi(rt.RenderOptions(128, 96, 2), [gpu_scene8], 2,
                                   on_frame=lambda f, a: got.__setitem__(f, a.copy()))
    assert np.array_equal(got[0], ref) and np.array_equal(got[1], ref)
