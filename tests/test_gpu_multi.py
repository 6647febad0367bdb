"""In-process multi-GPU frame (rt_render_frame_multi): interleaved rows + strided peer gather."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_multi_gpu_frame_matches_oracle(rt, oracle_scene8):
    n = rt.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 4)
    scenes = []
    for g in range(n):
        rt.set_device(g)
        scenes.append(rt.Scene())
    rt.set_device(0)
    for (w, h, spp) in [(256, 131, 2), (640, 360, 1)]:
        img, st = rt.Renderer.render_multi(rt.RenderOptions(w, h, spp), scenes, want_stats=True)
        ref, _ = oracle_scene8.render(w, h, spp)
        assert np.array_equal(img, ref)
        assert st.gpus == n


def test_single_scene_multi_call_is_a_plain_frame(rt, gpu_scene8, oracle_scene8):
    img = rt.Renderer.render_multi(rt.RenderOptions(128, 96, 2), [gpu_scene8])
    ref, _ = oracle_scene8.render(128, 96, 2)
    assert np.array_equal(img, ref)
