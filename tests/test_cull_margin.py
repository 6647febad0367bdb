"""The candidate-list kernels keep a leaf iff its exact f32 test *could* pass; that rests on one
number: EPS_DISC in rust-tracer_b200/csrc/rt_cull.cuh, the claimed worst-case rounding of the
reference's discriminant (primitive.rs:56-58)

    |disc_f32 - disc_exact| <= EPS_DISC * (v.v + r*r),

disc_f32 evaluated in the reference's order with its f32-normalised direction (vec.rs:78,87-95),
disc_exact in exact arithmetic with a truly unit direction.  This CPU test measures that error with
numpy (float32 ops round once each, like rustc's; float64 as "exact") over millions of random rays
and spheres of the scene's scales, so a change of EPS_DISC or of the op order shows up without a GPU."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def eps_disc():
    src = open(os.path.join(ROOT, "rust-tracer_b200", "csrc", "rt_cull.cuh")).read()
    return float(re.search(r"EPS_DISC\s*=\s*([0-9.eE+-]+)f", src).group(1))


def dot32(a, b):   # vec.rs:78: (x*x' + y*y') + z*z', one rounding per operation
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def normalize32(v):   # vec.rs:87-95: v * (1 / sqrt(v.v))
    k = F(1.0) / np.sqrt(dot32(v, v))
    return [v[0] * k, v[1] * k, v[2] * k]


def disc_error(v, raw_dir, r):
    """Relative error of the f32 discriminant, in units of (v.v + r*r)."""
    d = normalize32(raw_dir)
    b = dot32(v, d)
    disc32 = (b * b - dot32(v, v)) + r * r                       # primitive.rs:58, left to right
    v64 = [c.astype(np.float64) for c in v]
    d64 = [c.astype(np.float64) for c in raw_dir]
    n = np.sqrt(d64[0] ** 2 + d64[1] ** 2 + d64[2] ** 2)
    b64 = (v64[0] * d64[0] + v64[1] * d64[1] + v64[2] * d64[2]) / n
    vv = v64[0] ** 2 + v64[1] ** 2 + v64[2] ** 2
    rr = r.astype(np.float64) ** 2
    return np.abs(disc32.astype(np.float64) - (b64 * b64 - vv + rr)) / (vv + rr)


def random_cases(rng, n, width, height):
    # spheres of the pyramid's scales: leaf radii 2^-k and group bounds 3 * 2^-k, centres inside the root bound
    k = rng.integers(0, 10, n)
    r = (np.where(rng.random(n) < 0.5, 1.0, 3.0) * 2.0 ** (-k)).astype(F)
    c = [(rng.uniform(-3, 3, n)).astype(F), (rng.uniform(-4, 2, n)).astype(F), (rng.uniform(-3, 3, n)).astype(F)]
    return r, c


def test_primary_ray_discriminant_error_is_inside_the_cull_margin():
    rng = np.random.default_rng(20261017)
    n, worst = 1_000_000, 0.0
    for width, height in ((1024, 768), (3840, 2160), (7680, 4320)):
        r, c = random_cases(rng, n, width, height)
        eye = [F(0.0), F(0.0), F(-4.0)]
        v = [c[0] - eye[0], c[1] - eye[1], c[2] - eye[2]]          # primitive.rs:56, f32
        # render.rs:238-243: raw direction (x - W/2, (H - y) - H/2, W) at sub-sample positions k/4
        x = (rng.integers(0, width * 4, n) / 4.0).astype(F)
        y = (rng.integers(0, height * 4, n) / 4.0).astype(F)
        raw = [x - F(width) * F(0.5), (F(height) - y) - F(height) * F(0.5), np.full(n, F(width))]
        # half of the rays re-aimed at the sphere's silhouette, where the sign of disc decides hit or miss
        aim = rng.random(n) < 0.5
        t = F(width) / np.maximum(v[2], F(0.25))
        off = (r * t * rng.uniform(0.9, 1.1, n).astype(F))
        ang = rng.uniform(0, 2 * np.pi, n)
        raw[0] = np.where(aim, v[0] * t + off * np.cos(ang).astype(F), raw[0]).astype(F)
        raw[1] = np.where(aim, v[1] * t + off * np.sin(ang).astype(F), raw[1]).astype(F)
        worst = max(worst, float(disc_error(v, raw, r).max()))
    ulp = 2.0 ** -24
    assert worst <= eps_disc(), "discriminant error %.2f ulp exceeds EPS_DISC = %.2f ulp" % (worst / ulp, eps_disc() / ulp)
    assert worst >= 2 * ulp          # the measurement is live (not vacuously zero)


def test_shadow_ray_discriminant_error_is_inside_the_cull_margin():
    """Shadow rays (render.rs:199-207): origin on a sphere inside the scene, direction -light (f32-normalised)."""
    rng = np.random.default_rng(7)
    n = 2_000_000
    r, c = random_cases(rng, n, 0, 0)
    o = [(rng.uniform(-3, 3, n)).astype(F), (rng.uniform(-4, 2, n)).astype(F), (rng.uniform(-3, 3, n)).astype(F)]
    v = [c[0] - o[0], c[1] - o[1], c[2] - o[2]]
    light = normalize32([np.full(n, F(-1.0)), np.full(n, F(-3.0)), np.full(n, F(2.0))])
    to_light = [-light[0], -light[1], -light[2]]                # render.rs:206, already unit up to f32 rounding
    # the reference does not re-normalise: feed the f32 unit vector as the direction itself
    b = dot32(v, to_light)
    disc32 = (b * b - dot32(v, v)) + r * r
    v64 = [a.astype(np.float64) for a in v]
    l64 = np.array([-1.0, -3.0, 2.0]) / np.sqrt(14.0)
    b64 = -(v64[0] * l64[0] + v64[1] * l64[1] + v64[2] * l64[2])
    vv = v64[0] ** 2 + v64[1] ** 2 + v64[2] ** 2
    rr = r.astype(np.float64) ** 2
    err = np.abs(disc32.astype(np.float64) - (b64 * b64 - vv + rr)) / (vv + rr)
    assert float(err.max()) <= eps_disc(), "shadow discriminant error %.2f ulp" % (float(err.max()) / 2.0 ** -24)
