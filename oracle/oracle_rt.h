/*
 * oracle_rt.h -- CPU ORACLE for the rust-tracer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a from-scratch, op-order-faithful f32 restatement of the reference's
 * Rust path (/root/reference/src/rust/{vec,primitive,group,render}.rs).  It is
 * the checker the CUDA product is compared against; it is NOT part of the
 * product.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product library
 * (rust-tracer_b200/librtrace_b200.so) never links or calls anything here.
 *
 * Parity status: PINNED.  orc_render(1024,768,spp 4,level 8) reproduces the
 * reference's shipped golden image src/img/rtrace-output.png bit-exactly
 * (tests/test_oracle_golden.py; fixture tests/golden/rtrace_output_1024x768.json).
 *
 * Must be compiled with FP contraction OFF (-ffp-contraction=off), no fast-math,
 * SSE2 scalar f32 (x86-64 default): rustc never contracts a*b+c into an FMA.
 */
#ifndef ORACLE_RT_H
#define ORACLE_RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* Camera extension.  NULL camera == the reference's fixed camera
 * (render.rs:226-243): pos = scene.eye, dir = normalize(x-W/2,(H-y)-H/2,W). */
typedef struct orc_camera {
    float eye[3];
    float right[3];
    float up[3];
    float forward[3];
} orc_camera;

/* Work counters: the reference-work denominator of the roofline. */
typedef struct orc_counters {
    uint64_t primary_rays;   /* W*H*spp^2                                   */
    uint64_t shadow_rays;    /* samples with a hit and g < 0 (render.rs:194-203) */
    uint64_t primary_hits;
    uint64_t bound_tests;    /* distance_from_ray calls at group.rs:73      */
    uint64_t leaf_tests;     /* distance_from_ray calls at primitive.rs:78  */
    uint64_t disc_nonneg;    /* calls that reach the sqrt (primitive.rs:64) */
    uint64_t hit_updates;    /* accepted updates (primitive.rs:82-83)       */
} orc_counters;

typedef struct orc_ray { float pos[3]; float dir[3]; } orc_ray;
typedef struct orc_hit { float distance; float normal[3]; } orc_hit;

/* Scene::default generalised (render.rs:145-166): pyramid(level, origin, radius),
 * light = normalize(light_unnormalised), eye. */
orc_scene *orc_scene_create(uint32_t level, const float origin[3], float radius,
                            const float light_unnormalised[3], const float eye[3]);
/* The reference default: level 8, origin (0,-1,0), r 1, light (-1,-3,2), eye (0,0,-4). */
orc_scene *orc_scene_create_default(void);
/* Build an arbitrary group tree from a pre-order flat description (used for the
 * reference's group KAT, group.rs:118-151).  spheres4 = n x {cx,cy,cz,r};
 * skip[i] > i+1 marks node i as a group bound whose subtree ends at skip[i];
 * node 0 must be a group.  light is stored as given (already normalised). */
orc_scene *orc_scene_create_from_nodes(uint32_t n, const float *spheres4, const uint32_t *skip,
                                       const float light[3], const float eye[3]);
void orc_scene_destroy(orc_scene *s);

/* group.rs:95-109 TypedGroup::count -> (groups, items). */
void orc_scene_counts(const orc_scene *s, uint64_t *groups, uint64_t *items);
/* Pre-order flattening of the tree: group bound, then its children in order.
 * Returns node count; fills up to cap entries when the pointers are non-NULL. */
uint32_t orc_scene_flatten(const orc_scene *s, float *spheres4, uint32_t *skip, uint32_t cap);
void orc_scene_light(const orc_scene *s, float light[3]);
void orc_scene_eye(const orc_scene *s, float eye[3]);

/* primitive.rs:55-72 and :77-84 as free functions (KATs). */
float orc_sphere_distance_from_ray(const float center[3], float radius, const orc_ray *r);
void orc_sphere_intersect(const float center[3], float radius, orc_hit *h, const orc_ray *r);
/* vec.rs:87-95 */
void orc_vec_normalized(const float v[3], float out[3]);
float orc_vec_len(const float v[3]);

/* group.rs:72-83 closest-hit traversal for arbitrary rays; hits[i].distance=+inf on miss. */
void orc_trace_rays(const orc_scene *s, size_t n, const orc_ray *rays, orc_hit *hits);

/* render.rs:218-255 Renderer::render_region.  Region is [l,r) x [b,t), b = upper
 * image row.  rgba_out: (r-l)*(t-b)*4 bytes row-major from row b.  kinds_out
 * (optional): per-sample classification, (r-l)*(t-b)*spp*spp bytes, sample index
 * ssx*spp+ssy: 0 background, 1 hit facing away from light, 2 lit, 3 shadowed. */
void orc_render_region(const orc_scene *s, const orc_camera *cam,
                       uint32_t width, uint32_t height, uint32_t spp,
                       uint32_t l, uint32_t b, uint32_t r, uint32_t t,
                       uint8_t *rgba_out, uint8_t *kinds_out, orc_counters *ctr);

/* render.rs:260-310 Renderer::render: 64x64 buckets on nthreads threads, written
 * into the full W*H*4 frame.  Sizes need not be multiples of 64 (SURVEY F4:
 * edge buckets are clipped; for multiples of 64 the bucket set is the reference's). */
void orc_render(const orc_scene *s, const orc_camera *cam,
                uint32_t width, uint32_t height, uint32_t spp, uint32_t nthreads,
                uint8_t *rgba_out, orc_counters *ctr);

/* Render rows row_start, row_start+row_stride, ... (row_count rows) densely packed;
 * the multi-GPU interleaved partition's CPU twin. */
void orc_render_rows(const orc_scene *s, const orc_camera *cam,
                     uint32_t width, uint32_t height, uint32_t spp,
                     uint32_t row_start, uint32_t row_stride, uint32_t row_count,
                     uint32_t nthreads, uint8_t *rgba_out, orc_counters *ctr);

#ifdef __cplusplus
}
#endif
#endif
