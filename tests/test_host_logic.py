"""Host-side logic of the product library that needs no GPU: the C ABI loads and
exports every declared symbol, the scene flattening matches the oracle's tree
bit for bit, and argument validation / error reporting behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rtrace.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(rt):
    L = rt.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(L, name), "librtrace_b200.so does not export %s" % name
    assert sorted(rt.ABI_SYMBOLS) == syms


def test_no_torch_or_oracle_in_the_product_library():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "rust-tracer_b200", "librtrace_b200.so")], capture_output=True,
                         text=True).stdout
    assert "torch" not in out and "oracle" not in out and "python" not in out


def test_product_sources_never_reference_the_oracle():
    pkg = os.path.join(ROOT, "rust-tracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp", ".py")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_rt" not in text and "liboracle" not in text and "_oracle" not in text, f


@pytest.mark.parametrize("level", [2, 3, 5, 8, 9, 10])
def test_flatten_matches_oracle_tree(rt, oracle, level):
    sph, skip = rt.flatten_pyramid_host(level)
    osph, oskip = oracle.Scene(level=level).flatten()
    assert np.array_equal(skip, oskip)
    assert np.array_equal(sph.view(np.uint32), osph.view(np.uint32))  # bit-exact centres and radii
    groups = int((skip > np.arange(len(skip)) + 1).sum())
    assert (groups, len(skip) - groups) == oracle.Scene(level=level).counts()


def test_flatten_other_origin_and_radius(rt, oracle):
    sph, skip = rt.flatten_pyramid_host(6, origin=(1.0, -1.0, 0.0), radius=0.7)
    osph, oskip = oracle.Scene(level=6, origin=(1.0, -1.0, 0.0), radius=0.7).flatten()
    assert np.array_equal(skip, oskip) and np.array_equal(sph.view(np.uint32), osph.view(np.uint32))


def test_flatten_rejects_bad_levels(rt):
    for level in (0, 1, 13):
        with pytest.raises(rt.RtError) as e:
            rt.flatten_pyramid_host(level)
        assert e.value.code == rt.RT_ERR_INVALID
    assert "level" in str(e.value)


def test_skip_links_are_a_valid_preorder(rt):
    _, skip = rt.flatten_pyramid_host(7)
    n = len(skip)
    assert skip[0] == n
    stack = []
    for i in range(n):
        while stack and stack[-1] == i:
            stack.pop()
        limit = stack[-1] if stack else n
        assert i < skip[i] <= limit
        if skip[i] > i + 1:
            stack.append(int(skip[i]))


def test_compute_entry_points_fail_loudly_without_a_gpu(rt):
    if rt.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(rt.RtError) as e:
        rt.Scene()
    assert e.value.code == rt.RT_ERR_CUDA and "no CPU path" in str(e.value)
    with pytest.raises(rt.RtError):
        rt.measure_fp32_peak(0)
    with pytest.raises(rt.RtError):
        rt.PinnedBuffer(1024)


def test_invalid_variant_is_rejected(rt):
    with pytest.raises(rt.RtError):
        rt.set_variant(99)
    rt.set_variant(rt.VARIANT_AUTO)


def test_from_nodes_validation(rt):
    sph = np.zeros((3, 4), np.float32)
    for bad in ([2, 2, 3], [3, 1, 3], [3, 2, 5], [3, 2, 2]):
        with pytest.raises(rt.RtError) as e:
            rt.Scene.from_nodes(sph, np.array(bad, np.uint32), (0, -1, 0), (0, 0, -4))
        assert e.value.code == rt.RT_ERR_INVALID


@pytest.mark.parametrize("level", [3, 6, 8])
def test_every_leaf_lies_two_leaf_radii_inside_each_ancestor_bound(rt, level):
    """The candidate-list kernels replace the reference's pruned walk (group.rs:72-83) by "nearest of the leaves
    whose exact test passes".  That is the same answer only if a bound can never prune a leaf that would have
    won, i.e. every leaf sphere lies strictly inside each ancestor's bounding sphere with room to spare
    against f32 noise: the pyramid (group.rs:28-56) leaves 2 x the smallest leaf radius."""
    sph, skip = rt.flatten_pyramid_host(level)
    n = len(skip)
    leaf = skip == np.arange(n) + 1
    r_min = float(sph[leaf, 3].min())
    c64 = sph[:, :3].astype(np.float64)
    stack, slack = [], np.inf
    for i in range(n):
        while stack and skip[stack[-1]] <= i:
            stack.pop()
        if not leaf[i]:
            stack.append(i)
            continue
        anc = np.array(stack)
        d = np.linalg.norm(c64[anc] - c64[i], axis=1) + float(sph[i, 3])
        slack = min(slack, float((sph[anc, 3] - d).min()))
    assert slack >= 1.99 * r_min, (slack, r_min)


def test_frame_queue_counter_is_an_atomic_fetch_add():
    """rt_atomic_fetch_add_u64: the frame queue of a multi-process sweep (a counter in shared memory); host-only."""
    import ctypes
    import threading
    import rtrace_b200 as rt
    word = ctypes.c_uint64(0)
    addr = ctypes.addressof(word)
    seen = [[] for _ in range(4)]

    def worker(k):
        for _ in range(2000):
            seen[k].append(rt.atomic_fetch_add_u64(addr, 1))
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert word.value == 8000
    assert sorted(v for s in seen for v in s) == list(range(8000))   # every ticket handed out exactly once
