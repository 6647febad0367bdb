"""Command-line conformance of target/release/rtrace with the reference binary (src/rust/main.rs).

The argument / environment / exit-code behaviour needs no GPU; the rendering runs are GPU tests and
compare the written file with the reference's golden `make image` output."""
import hashlib
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "target", "release", "rtrace")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "rtrace_output_1024x768.json")))


@pytest.fixture(scope="module")
def rtrace():
    if not os.path.exists(BIN):
        subprocess.check_call(["make", "-s", "-C", ROOT, "rtrace"])
    return BIN


def run(rtrace, *args, env=None, cwd=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([rtrace] + list(args), capture_output=True, env=e, cwd=cwd, timeout=120)


def test_missing_output_is_a_usage_error(rtrace):
    r = run(rtrace)
    assert r.returncode == 1 and b"<output>" in r.stderr and r.stdout == b""


def test_bad_extension_prints_hint_on_stdout_and_exits_zero(rtrace, tmp_path):
    # main.rs:66-71
    r = run(rtrace, "picture.png", cwd=str(tmp_path))
    assert r.returncode == 0
    assert r.stdout == b"Output file 'picture.png' must have the tga extension, e.g. picture.tga\n"
    assert not (tmp_path / "picture.png").exists()
    r = run(rtrace, "noext", cwd=str(tmp_path))
    assert r.returncode == 0 and b"noext.tga" in r.stdout


def test_unparsable_numbers_panic_with_101(rtrace, tmp_path):
    # main.rs:56,79-81: `.parse().unwrap()`
    for flag in ("--width=abc", "--height=-3", "--samples-per-pixel=1.5", "--width=70000", "--num-cores=x"):
        r = run(rtrace, flag, "o.tga", cwd=str(tmp_path))
        assert r.returncode == 101, flag
        assert b"panicked" in r.stderr


def test_extension_flags_are_parsed_like_the_reference_flags(rtrace, tmp_path):
    """--preview / --level / --frames / --gpus values go through the same `.parse().unwrap()` rule
    (main.rs:79-81): a bad number panics with 101 before any GPU work; --help lists the additions."""
    for flag in ("--preview=abc", "--preview=5000", "--level=x", "--frames=-1", "--gpus=two"):
        r = run(rtrace, flag, "o.tga", cwd=str(tmp_path))
        assert r.returncode == 101 and b"panicked" in r.stderr, flag
    assert run(rtrace, "--preview", cwd=str(tmp_path)).returncode == 1   # flag without its value: usage error
    h = run(rtrace, "--help").stdout
    assert b"--preview <N>" in h and b"--buckets" in h


def test_unknown_flag_and_extra_positional_are_usage_errors(rtrace, tmp_path):
    assert run(rtrace, "--bogus", "o.tga", cwd=str(tmp_path)).returncode == 1
    assert run(rtrace, "a.tga", "b.tga", cwd=str(tmp_path)).returncode == 1
    assert run(rtrace, "--width", cwd=str(tmp_path)).returncode == 1
    assert run(rtrace, "", cwd=str(tmp_path)).returncode == 1   # empty_values(false), main.rs:53


def test_help_and_version(rtrace):
    r = run(rtrace, "--help")
    assert r.returncode == 0 and b"--samples-per-pixel" in r.stdout and b"--num-cores" in r.stdout
    r = run(rtrace, "--version")
    assert r.returncode == 0 and r.stdout.strip() == b"rtrace 0.2.0"   # main.rs:33


def test_makefile_keeps_the_reference_targets():
    mk = open(os.path.join(ROOT, "Makefile")).read()
    for target in ("all: rtrace", "image: rtrace"):
        assert target in mk
    # the exact command line of the reference `make image` (Makefile:7)
    assert "time ./target/release/rtrace --samples-per-pixel=4 --width=1024 --height=768 out.tga" in mk


@pytest.mark.gpu
def test_make_image_file_equals_reference_golden(rtrace, tmp_path):
    """`make image`: out.tga is a binary PPM whose bytes equal the reference's shipped image."""
    r = run(rtrace, "--samples-per-pixel=4", "--width=1024", "--height=768", "out.tga", cwd=str(tmp_path),
            env={"RTRACEMAXPROCS": "4"})
    assert r.returncode == 0, r.stderr
    data = (tmp_path / "out.tga").read_bytes()
    hdr = b"P6\n1024 768\n255\n"
    assert data.startswith(hdr) and len(data) == len(hdr) + 1024 * 768 * 3 == 2359312
    assert hashlib.sha256(data).hexdigest() == GOLD["ppm_sha256"]


@pytest.mark.gpu
def test_stdout_sink_and_flag_forms(rtrace, tmp_path):
    a = run(rtrace, "--width", "64", "--height=128", "--samples-per-pixel", "2", "-")
    hdr = b"P6\n64 128\n255\n"
    assert a.returncode == 0 and a.stdout.startswith(hdr) and len(a.stdout) == len(hdr) + 64 * 128 * 3
    b = run(rtrace, "--width=64", "--height=128", "--samples-per-pixel=2", "--num-cores=3", "x.tga", cwd=str(tmp_path))
    assert b.returncode == 0 and (tmp_path / "x.tga").read_bytes() == a.stdout


@pytest.mark.gpu
def test_sizes_the_reference_rejects_and_extensions(rtrace, tmp_path):
    # 100x60 is not a multiple of 64: the reference asserts (render.rs:265-266); here it renders
    r = run(rtrace, "--width=100", "--height=60", "--level=5", "--stats", "e.tga", cwd=str(tmp_path))
    assert r.returncode == 0 and b"rtrace-b200:" in r.stderr
    assert (tmp_path / "e.tga").read_bytes().startswith(b"P6\n100 60\n255\n")
    # orbit sweep: one file per frame, frame 0 is the reference camera
    r = run(rtrace, "--width=64", "--height=64", "--frames=3", "s.tga", cwd=str(tmp_path))
    assert r.returncode == 0
    frames = [(tmp_path / ("s.%04d.tga" % f)).read_bytes() for f in range(3)]
    one = run(rtrace, "--width=64", "--height=64", "-")
    assert frames[0] == one.stdout and frames[1] != frames[0]


@pytest.mark.gpu
def test_bucket_schedule_writes_the_same_file(rtrace, tmp_path):
    """--buckets: the reference's own schedule (render.rs:265-309), 64x64 buckets through render_region and
    the writer seam one by one; the file equals the whole-frame path's, and sizes that are not multiples of
    64 panic as in the reference (render.rs:265-266)."""
    a = run(rtrace, "--width=256", "--height=192", "--samples-per-pixel=2", "-")
    b = run(rtrace, "--width=256", "--height=192", "--samples-per-pixel=2", "--buckets", "b.tga", cwd=str(tmp_path))
    assert a.returncode == 0 and b.returncode == 0, b.stderr
    assert (tmp_path / "b.tga").read_bytes() == a.stdout
    r = run(rtrace, "--width=100", "--height=64", "--buckets", "c.tga", cwd=str(tmp_path))
    assert r.returncode == 101 and b"BUCKET_SIZE" in r.stderr


@pytest.mark.gpu
def test_preview_flag(rtrace, tmp_path):
    """--preview N (README.md:42-48, extension): blocks of NxN pixels share the colour of their first pixel."""
    w, h, n = 96, 64, 8
    full = run(rtrace, "--width=%d" % w, "--height=%d" % h, "-")
    prev = run(rtrace, "--width=%d" % w, "--height=%d" % h, "--preview=%d" % n, "-")
    assert full.returncode == 0 and prev.returncode == 0, prev.stderr
    hdr = len(b"P6\n96 64\n255\n")
    f, p = full.stdout[hdr:], prev.stdout[hdr:]
    assert len(p) == w * h * 3
    for y in range(h):
        for x in range(w):
            a = ((y // n * n) * w + (x // n * n)) * 3
            assert p[(y * w + x) * 3:(y * w + x) * 3 + 3] == f[a:a + 3]


def test_format_flag_is_validated(rtrace, tmp_path):
    assert run(rtrace, "--format=bmp", "o.tga", cwd=str(tmp_path)).returncode == 1


@pytest.mark.gpu
def test_real_tga_output_holds_the_same_pixels(rtrace, tmp_path):
    """--format tga (extension): 18-byte header, 24-bit BGR, bottom-up (gotrace.go:248-280)."""
    w, h = 96, 64
    ppm = run(rtrace, "--width=%d" % w, "--height=%d" % h, "--samples-per-pixel=2", "-")
    r = run(rtrace, "--width=%d" % w, "--height=%d" % h, "--samples-per-pixel=2", "--format=tga", "t.tga", cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    data = (tmp_path / "t.tga").read_bytes()
    assert len(data) == 18 + w * h * 3
    assert data[:18] == bytes([0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, w & 255, w >> 8, h & 255, h >> 8, 24, 0])
    rgb = ppm.stdout[len(b"P6\n96 64\n255\n"):]
    rows = [rgb[y * w * 3:(y + 1) * w * 3] for y in range(h)]
    expect = b"".join(bytes(b for x in range(w) for b in (row[3 * x + 2], row[3 * x + 1], row[3 * x])) for row in reversed(rows))
    assert data[18:] == expect
    # the sweep path writes the same format
    r = run(rtrace, "--width=%d" % w, "--height=%d" % h, "--samples-per-pixel=2", "--format=tga", "--frames=2", "s.tga", cwd=str(tmp_path))
    assert r.returncode == 0 and (tmp_path / "s.0000.tga").read_bytes() == data


SELFTEST = os.path.join(ROOT, "target", "release", "rtrace_selftest")


def selftest(rtrace, *args):
    assert os.path.exists(SELFTEST), "make rtrace builds target/release/rtrace_selftest"
    return subprocess.run([SELFTEST] + list(args), capture_output=True, timeout=120)


def test_image_region_kat(rtrace):
    """render.rs:483-499 `image_region`, value for value, on the C++ mirror of ImageRegion (no GPU needed)."""
    r = selftest(rtrace, "image_region")
    assert r.returncode == 0 and b"ok image_region" in r.stdout, r.stderr


@pytest.mark.gpu
def test_writer_seam_kat_basic_rendering(rtrace):
    """render.rs:466-481 `basic_rendering`: 64x128, 2x2 samples, a counting DummyWriter sees begin() and exactly
    two bucket writes when the frame goes through the reference's 64x64 schedule (Renderer::render_buckets)."""
    r = selftest(rtrace, "basic_rendering")
    assert r.returncode == 0 and b"ok basic_rendering" in r.stdout, r.stderr


@pytest.mark.gpu
def test_progressive_rewrite_once_per_second(rtrace, tmp_path):
    """render.rs:426-432: a file sink rewrites the whole image on the first buffer and then at most once per
    second; the drop writes the final image (render.rs:331-335); a non-file sink writes only on drop."""
    r = selftest(rtrace, "progressive", str(tmp_path / "p.tga"))
    assert r.returncode == 0 and b"ok progressive" in r.stdout, r.stderr


@pytest.mark.gpu
def test_sink_exception_does_not_cross_the_c_boundary(rtrace):
    r = selftest(rtrace, "sink_error")
    assert r.returncode == 0 and b"ok sink_error" in r.stdout, r.stderr


def _gpus():
    import rtrace_b200 as rt
    return rt.device_count()


@pytest.mark.gpu
def test_two_gpu_frame_and_sweep_write_the_same_files(rtrace, tmp_path):
    """`--gpus 2`: one frame split into interleaved row blocks (rt_render_frame_multi) and a sweep sharded by
    frame (rt_render_sweep_multi) write exactly the files one GPU writes."""
    if _gpus() < 2:
        pytest.skip("needs at least 2 GPUs")
    one = run(rtrace, "--width=320", "--height=200", "--samples-per-pixel=2", "-")
    two = run(rtrace, "--width=320", "--height=200", "--samples-per-pixel=2", "--gpus=2", "--stats", "-")
    assert one.returncode == 0 and two.returncode == 0, two.stderr
    assert one.stdout == two.stdout and b"on 2 GPU(s)" in two.stderr
    a = run(rtrace, "--width=160", "--height=96", "--frames=5", "a.tga", cwd=str(tmp_path))
    b = run(rtrace, "--width=160", "--height=96", "--frames=5", "--gpus=2", "b.tga", cwd=str(tmp_path),
            env={"RTRACE_GPUS": "1"})   # the flag overrides the environment
    assert a.returncode == 0 and b.returncode == 0, b.stderr
    for f in range(5):
        assert (tmp_path / ("a.%04d.tga" % f)).read_bytes() == (tmp_path / ("b.%04d.tga" % f)).read_bytes(), f
    # `make image` on two GPUs is still the reference's golden file
    r = run(rtrace, "--samples-per-pixel=4", "--width=1024", "--height=768", "out.tga", cwd=str(tmp_path),
            env={"RTRACE_GPUS": "2"})
    assert r.returncode == 0, r.stderr
    assert hashlib.sha256((tmp_path / "out.tga").read_bytes()).hexdigest() == GOLD["ppm_sha256"]
