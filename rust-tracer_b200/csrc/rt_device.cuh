// rt_device.cuh -- strict-f32 device math and the kernel parameter block.
//
// Parity rule: the reference is Rust f32 with no FMA contraction, so
// every arithmetic node on the parity-critical path is a single correctly
// rounded IEEE operation issued through an *_rn intrinsic (nvcc never fuses
// those), in the reference's evaluation order.  The file is additionally
// compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

#define RT_DEV __device__ __forceinline__

RT_DEV float fadd(float a, float b) { return __fadd_rn(a, b); }
RT_DEV float fsub(float a, float b) { return __fsub_rn(a, b); }
RT_DEV float fmul(float a, float b) { return __fmul_rn(a, b); }
RT_DEV float fsqrt(float a) { return __fsqrt_rn(a); }
RT_DEV float frecip(float a) { return __frcp_rn(a); }  // == IEEE 1.0f / a (f32::recip)
RT_DEV float fdiv(float a, float b) { return __fdiv_rn(a, b); }

struct V3 {
    float x, y, z;
};
RT_DEV V3 v3(float x, float y, float z) { return V3{x, y, z}; }
// vec.rs:20-26, :33-39, :57-63
RT_DEV V3 vadd(V3 a, V3 b) { return v3(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z)); }
RT_DEV V3 vsub(V3 a, V3 b) { return v3(fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)); }
RT_DEV V3 vmulf(V3 a, float m) { return v3(fmul(a.x, m), fmul(a.y, m), fmul(a.z, m)); }
// vec.rs:77-79: (x*x' + y*y') + z*z', left to right
RT_DEV float vdot(V3 a, V3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
// vec.rs:87-95: v * (1/sqrt(v.v))
RT_DEV V3 vnormalized(V3 a) { return vmulf(a, frecip(fsqrt(vdot(a, a)))); }

#define RT_INF __int_as_float(0x7f800000)

// primitive.rs:55-72 Sphere::distance_from_ray
RT_DEV float sphere_distance(float4 s, V3 o, V3 d) {
    V3 v = vsub(v3(s.x, s.y, s.z), o);
    float b = vdot(v, d);
    float disc = fadd(fsub(fmul(b, b), vdot(v, v)), fmul(s.w, s.w));
    if (disc < 0.0f) return RT_INF;
    float sq = fsqrt(disc);
    float t2 = fadd(b, sq);
    if (t2 < 0.0f) return RT_INF;
    float t1 = fsub(b, sq);
    return t1 > 0.0f ? t1 : t2;
}

// "is distance_from_ray finite" without the final select: the any-hit form used
// for shadow rays (render.rs:202-208 reads only has_missed()).
RT_DEV bool sphere_hit_any(float4 s, V3 o, V3 d) {
    V3 v = vsub(v3(s.x, s.y, s.z), o);
    float b = vdot(v, d);
    float disc = fadd(fsub(fmul(b, b), vdot(v, v)), fmul(s.w, s.w));
    if (disc < 0.0f) return false;
    float t2 = fadd(b, fsqrt(disc));
    return !(t2 < 0.0f);
}

// primitive.rs:83: normalize(ray.pos + (ray.dir * distance - center))
RT_DEV V3 hit_normal(float4 s, V3 o, V3 d, float dist) {
    return vnormalized(vadd(o, vsub(vmulf(d, dist), v3(s.x, s.y, s.z))));
}

// render.rs:96-103: trunc(0.5 + 255 v), > 255 -> 255, Rust `as u8` saturates (NaN -> 0)
RT_DEV uint32_t scale_u8(float v) {
    float r = fadd(0.5f, fmul(255.0f, v));
    if (r > 255.0f) return 255u;
    if (!(r > 0.0f)) return 0u;
    return (uint32_t)r;  // cvt.rzi
}

enum SampleKind : uint8_t { K_BACKGROUND = 0, K_AWAY = 1, K_LIT = 2, K_SHADOWED = 3 };

// Kernel parameter block (passed by value: lives in the constant bank).
struct RenderParams {
    const float4 *sph;     // n x {cx,cy,cz,r}, pre-order
    const uint32_t *skip;  // n x next-node-when-pruned; leaf: i+1
    uint32_t n_nodes;
    uint32_t level;  // pyramid level (regular tree: child offsets follow from the depth); 0 = hand-built tree
    float leaf_rmin;  // smallest leaf radius (regular pyramid)
    float scene_center[3];  // centre of the root bound (host-side heuristics only)
    float scene_radius;     // radius of the root bound
    float one;             // 1.0f, deliberately a runtime value (rt_pack.cuh)
    uint32_t tile_stride;  // TILE variant: warp w renders tile (w * tile_stride) mod n_tiles
    float eye[3];
    float light[3];  // normalised directional light (render.rs:154-159)
    float lframe[6];  // e1, e2: orthonormal pair perpendicular to the light (PHASED shadow pre-filter; not parity arithmetic)
    float basis[9];  // right, up, forward (camera extension)
    int has_basis;
    uint32_t width, height, spp;
    uint32_t row_start, row_stride, row_count;  // image rows rendered by this launch
    uint32_t row_block_shift;                   // rows come in blocks of 2^shift consecutive image rows
    int out_abs;                                // rows are stored at their IMAGE row (out = base of a whole frame)
    uint8_t *out;                               // RGBA8, row j at out + j*pitch
    size_t pitch;
    // PHASED variant scratch (L2-resident intermediates between the four launches)
    uint4 *pool;            // candidate chunks, 16-byte units
    uint32_t pool_cap;      // units
    uint32_t *pool_count;   // units used (atomic)
    uint4 *tile_hdr;        // per cull tile {primary chain, shadow chain, tmin bits, tmax bits}
    uint32_t *winner;       // per sample: index of the closest leaf
    uint8_t *kinds;                    // optional per-sample classification
    unsigned long long *ray_counters;  // optional {primary_hits, shadow_rays}
    // Column window of Renderer::render_region (render.rs:218-255, a bucket): pixels col_start ..
    // col_start + col_count - 1 of each row, stored from byte 0 of the output row.  col_count 0 = the
    // whole width.  Only the per-lane kernels (LANE / WARP) take a window; the API routes it there.
    uint32_t col_start, col_count;
    // Undersampled preview (rt_render_preview): lane (lx, j) traces the one ray of pixel (lx*px_step, j*px_step)
    // and stores its colour into the whole px_step x px_step block.  Read by the PREVIEW instantiation only.
    uint32_t px_step;
};

// Local row j of this launch -> image row: blocks of 2^shift consecutive rows, row_stride apart
// (shift 0: the plain interleave row_start + j * row_stride).
RT_DEV uint32_t image_row(const RenderParams &p, uint32_t j) {
    return p.row_start + (j >> p.row_block_shift) * p.row_stride + (j & ((1u << p.row_block_shift) - 1u));
}
// Row of the output buffer that local row j is stored in.
RT_DEV uint32_t out_row(const RenderParams &p, uint32_t j) { return p.out_abs ? image_row(p, j) : j; }

// Shading constants, render.rs:172-186 (single f32 operations, as rustc const-evaluates them).
struct ShadeConsts {
    V3 object, background, ambient;
};
RT_DEV ShadeConsts shade_consts() {
    ShadeConsts c;
    c.object = v3(fdiv(174.0f, 255.0f), fdiv(49.0f, 255.0f), fdiv(49.0f, 255.0f));
    c.background = v3(fdiv(34.0f, 255.0f), fdiv(10.0f, 255.0f), fdiv(10.0f, 255.0f));
    c.ambient = v3(fmul(c.background.x, 0.8f), fmul(c.background.y, 0.8f), fmul(c.background.z, 0.8f));
    return c;
}

// render.rs:238-243 ray generation for sample (ssx, ssy) of pixel (x, y).
RT_DEV V3 primary_dir(const RenderParams &p, uint32_t x, uint32_t y, uint32_t ssx, uint32_t ssy) {
    float ssf = (float)p.spp;
    float width = (float)p.width, height = (float)p.height;
    float xres = fadd((float)x, fdiv((float)ssx, ssf));
    float yres = fadd((float)y, fdiv((float)ssy, ssf));
    V3 d;
    d.x = fsub(xres, fmul(width, 0.5f));                  // width / 2.0 (exact either way)
    d.y = fsub(fsub(height, yres), fmul(height, 0.5f));  // (height - yres) - height / 2.0
    d.z = width;
    if (p.has_basis) {
        V3 w;
        w.x = fadd(fadd(fmul(p.basis[0], d.x), fmul(p.basis[3], d.y)), fmul(p.basis[6], d.z));
        w.y = fadd(fadd(fmul(p.basis[1], d.x), fmul(p.basis[4], d.y)), fmul(p.basis[7], d.z));
        w.z = fadd(fadd(fmul(p.basis[2], d.x), fmul(p.basis[5], d.y)), fmul(p.basis[8], d.z));
        d = w;
    }
    return vnormalized(d);
}

// ---------------------------------------------------------------------------
// LANE traversal: per-lane stackless walk, exactly the reference recursion.
// ---------------------------------------------------------------------------
template <bool ANY>
RT_DEV void lane_traverse(const float4 *__restrict__ sph, const uint32_t *__restrict__ skip, uint32_t n, V3 o, V3 d,
                          float &hitd, uint32_t &hit_idx) {
    uint32_t i = 0;
    while (i < n) {
        float4 s = __ldg(&sph[i]);
        uint32_t sk = __ldg(&skip[i]);
        if (sk > i + 1) {  // group bound: group.rs:73-75
            bool enter;
            if (ANY)
                enter = sphere_hit_any(s, o, d);
            else
                enter = !(sphere_distance(s, o, d) >= hitd);
            i = enter ? i + 1 : sk;
        } else {  // leaf: primitive.rs:77-84
            if (ANY) {
                if (sphere_hit_any(s, o, d)) {
                    hitd = 0.0f;
                    return;
                }
            } else {
                float dist = sphere_distance(s, o, d);
                if (!(dist >= hitd)) {
                    hitd = dist;
                    hit_idx = i;
                }
            }
            i = i + 1;
        }
    }
}


}  // namespace rt
