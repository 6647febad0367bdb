// rt_pack.cuh -- the strict-f32 arithmetic of rt_device.cuh, two rays at a time.
//
// sm_100 has packed f32x2 instructions (FMUL2 / FADD2 / FFMA2, PTX mul/add/fma.rn.f32x2).
// Each component is rounded to nearest exactly like the scalar instruction, so a pair of
// rays can share every instruction of the reference's arithmetic without changing a bit.
// Measured on B200 (rt_microbench_fp32): scalar FMUL+FADD chains reach 36.2 TFLOP/s, the
// packed FMUL2+FADD2 chains 72.4 TFLOP/s -- the FFMA rate.  The parity rule
// forbids fusing b*b - v.v + r*r into FMAs; packing gives that throughput back.
//
// ptxas (CUDA 12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 at every -O level above 0,
// regardless of the .rn qualifiers and of -fmad=false (verified in SASS) -- which would change the
// rounding.  Every packed ADDITION is therefore issued as FFMA2(a, ONE, b) with ONE = 1.0f read
// from the kernel parameters (unknown to the compiler): a * 1 + b rounds exactly like a + b, and a
// product feeding it stays a separate FMUL2.  rt_selftest_math checks the packed functions
// against the scalar ones bit for bit.
#pragma once
#include "rt_cull.cuh"

namespace rt {

typedef float2 F2;
RT_DEV F2 f2(float a, float b) { return make_float2(a, b); }
RT_DEV F2 f2s(float a) { return make_float2(a, a); }
RT_DEV F2 f2neg(F2 a) { return make_float2(-a.x, -a.y); }
typedef F2 ONE2;                                                            // (1.0f, 1.0f), a runtime value
RT_DEV F2 f2add(ONE2 one, F2 a, F2 b) { return __ffma2_rn(a, one, b); }     // a * 1 + b == a + b exactly
RT_DEV F2 f2sub(ONE2 one, F2 a, F2 b) { return __ffma2_rn(a, one, f2neg(b)); }  // a - b == a + (-b) exactly
RT_DEV F2 f2mul(F2 a, F2 b) { return __fmul2_rn(a, b); }
RT_DEV F2 f2fma(F2 a, F2 b, F2 c) { return __ffma2_rn(a, b, c); }

struct V3x2 {
    F2 x, y, z;
};
RT_DEV V3x2 v3x2(V3 a, V3 b) { return V3x2{f2(a.x, b.x), f2(a.y, b.y), f2(a.z, b.z)}; }
RT_DEV V3x2 v3x2s(V3 a) { return V3x2{f2s(a.x), f2s(a.y), f2s(a.z)}; }
RT_DEV V3 lo(V3x2 a) { return v3(a.x.x, a.y.x, a.z.x); }
RT_DEV V3 hi(V3x2 a) { return v3(a.x.y, a.y.y, a.z.y); }
RT_DEV V3x2 vadd2(ONE2 k, V3x2 a, V3x2 b) { return V3x2{f2add(k, a.x, b.x), f2add(k, a.y, b.y), f2add(k, a.z, b.z)}; }
RT_DEV V3x2 vsub2(ONE2 k, V3x2 a, V3x2 b) { return V3x2{f2sub(k, a.x, b.x), f2sub(k, a.y, b.y), f2sub(k, a.z, b.z)}; }
RT_DEV V3x2 vmulf2(V3x2 a, F2 m) { return V3x2{f2mul(a.x, m), f2mul(a.y, m), f2mul(a.z, m)}; }
// vec.rs:77-79: (x*x' + y*y') + z*z'
RT_DEV F2 vdot2(ONE2 k, V3x2 a, V3x2 b) { return f2add(k, f2add(k, f2mul(a.x, b.x), f2mul(a.y, b.y)), f2mul(a.z, b.z)); }

// fsqrt_nr / frecip_nr (rt_cull.cuh) on both components: the MUFU seeds are scalar, the
// Newton steps packed.
RT_DEV F2 fsqrt_nr2(F2 x) {
    float y0, y1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x.x));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(x.y));
    const F2 y = f2(y0, y1);
    F2 s = f2mul(x, y);
    const F2 h = f2mul(y, f2s(0.5f));
    const F2 e = f2fma(f2neg(s), s, x);
    s = f2fma(e, h, s);
    return f2(x.x == 0.0f ? x.x : s.x, x.y == 0.0f ? x.y : s.y);
}
RT_DEV F2 frecip_nr2(F2 x) {
    float r0, r1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(x.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(x.y));
    const F2 r = f2(r0, r1);
    const F2 e = f2fma(x, r, f2s(-1.0f));
    return f2fma(r, f2neg(e), r);
}
// fsqrt_nr2 for arguments that are never zero: the squared length of a ray direction (>= width^2) or of a hit
// point's offset from its sphere's centre (~r^2).  Same Newton step, same bits; the x == 0 selects (two compares and
// two selects per pair) are dropped.  A lane without a hit normalises a dummy vector whose result is discarded.
RT_DEV F2 fsqrt_nr2_nonzero(F2 x) {
    float y0, y1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x.x));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(x.y));
    const F2 y = f2(y0, y1);
    const F2 s = f2mul(x, y);
    const F2 h = f2mul(y, f2s(0.5f));
    const F2 e = f2fma(f2neg(s), s, x);
    return f2fma(e, h, s);
}
// vec.rs:87-95.  NONZERO: the caller knows the vector cannot be the zero vector (measured on B200: the supersampled K2 / K4
// gain 3 % from the shorter sqrt, the one-sample-per-pixel kernels LOSE 1.7 % -- a scheduling accident of their
// instantiation -- so those keep the guarded form; the bits are the same either way).
template <bool NONZERO = false>
RT_DEV V3x2 vnormalized2(ONE2 k, V3x2 a) {
    const F2 n2 = vdot2(k, a, a);
    return vmulf2(a, frecip_nr2(NONZERO ? fsqrt_nr2_nonzero(n2) : fsqrt_nr2(n2)));
}

// primitive.rs:55-72 for two rays against one candidate given as broadcast pairs
// v = c - eye, nvv = -(v.v), rr = r*r.  Branch-free: a miss yields +inf.
RT_DEV F2 primary_distance2(ONE2 k, V3x2 v, F2 nvv, F2 rr, V3x2 d) {
    const F2 b = vdot2(k, v, d);
    const F2 disc = f2add(k, f2add(k, f2mul(b, b), nvv), rr);  // (b*b - v.v) + r*r
    const F2 sq = fsqrt_nr2(disc);
    const F2 t2 = f2add(k, b, sq);
    const F2 t1 = f2sub(k, b, sq);
    F2 out;
    out.x = (disc.x < 0.0f || t2.x < 0.0f) ? RT_INF : (t1.x > 0.0f ? t1.x : t2.x);
    out.y = (disc.y < 0.0f || t2.y < 0.0f) ? RT_INF : (t1.y > 0.0f ? t1.y : t2.y);
    return out;
}

// render.rs:238-243 for two slots of one lane (sub-sample offsets folded at compile time).
template <int SPP>
RT_DEV V3x2 slot_dir2(ONE2 k, const RenderParams &p, uint32_t x0, uint32_t y0, int smp0, uint32_t x1, uint32_t y1, int smp1) {
    constexpr float off0 = 0.0f / SPP, off1 = 1.0f / SPP, off2 = 2.0f / SPP, off3 = 3.0f / SPP;
    auto off = [&](int k) {
#ifndef RT_NO_SUBOFF_TABLE
        if constexpr (SPP >= 2) return subsample_offset_table<SPP>(k);
#endif
        if constexpr (SPP <= 4) return k == 0 ? off0 : k == 1 ? off1 : k == 2 ? off2 : off3;
        else return subsample_offset_wide<SPP>(k);
    };
    const float width = (float)p.width, height = (float)p.height;
    const F2 xres = f2add(k, f2((float)x0, (float)x1), f2(off(smp0 / SPP), off(smp1 / SPP)));
    const F2 yres = f2add(k, f2((float)y0, (float)y1), f2(off(smp0 % SPP), off(smp1 % SPP)));
    V3x2 d;
    d.x = f2sub(k, xres, f2s(fmul(width, 0.5f)));
    d.y = f2sub(k, f2sub(k, f2s(height), yres), f2s(fmul(height, 0.5f)));
    d.z = f2s(width);
    if (p.has_basis) {
        V3x2 w;
        w.x = f2add(k, f2add(k, f2mul(f2s(p.basis[0]), d.x), f2mul(f2s(p.basis[3]), d.y)), f2mul(f2s(p.basis[6]), d.z));
        w.y = f2add(k, f2add(k, f2mul(f2s(p.basis[1]), d.x), f2mul(f2s(p.basis[4]), d.y)), f2mul(f2s(p.basis[7]), d.z));
        w.z = f2add(k, f2add(k, f2mul(f2s(p.basis[2]), d.x), f2mul(f2s(p.basis[5]), d.y)), f2mul(f2s(p.basis[8]), d.z));
        d = w;
    }
    return vnormalized2<(SPP > 1)>(k, d);
}

}  // namespace rt
