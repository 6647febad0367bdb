/*
 * oracle_rt.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see oracle_rt.h).
 *
 * Every function cites the reference lines it restates.  The arithmetic is
 * written one IEEE f32 operation per expression node, in the reference's
 * evaluation order, and the file must be compiled with -ffp-contract=off.
 */
#include "oracle_rt.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

typedef float RFloat;  // vec.rs:6
const RFloat INF = std::numeric_limits<float>::infinity();

// ---- vec.rs:8-96 ---------------------------------------------------------
struct Vector {
    RFloat x, y, z;
};
inline Vector add(Vector a, Vector b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }  // vec.rs:20-26
inline Vector sub(Vector a, Vector b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }  // vec.rs:33-39
inline Vector mulfed(Vector a, RFloat m) { return {a.x * m, a.y * m, a.z * m}; }      // vec.rs:57-63
inline RFloat dot(Vector a, Vector b) { return a.x * b.x + a.y * b.y + a.z * b.z; }   // vec.rs:77-79 (left to right)
inline RFloat len(Vector a) { return sqrtf(dot(a, a)); }                              // vec.rs:82-84
inline RFloat recip(RFloat v) { return 1.0f / v; }                                    // f32::recip
inline Vector normalized(Vector a) { return mulfed(a, recip(len(a))); }               // vec.rs:87-95

// ---- primitive.rs:9-36 ---------------------------------------------------
struct Ray {
    Vector pos, dir;
};
struct Hit {
    RFloat distance;
    Vector pos;
    bool has_missed() const { return distance == INF; }  // primitive.rs:29-31
};
struct Sphere {
    Vector center;
    RFloat radius;
};

struct Counters {
    uint64_t bound_tests = 0, leaf_tests = 0, disc_nonneg = 0, hit_updates = 0;
};

// primitive.rs:55-72 Sphere::distance_from_ray
inline RFloat distance_from_ray(const Sphere &s, const Ray &r, Counters &k) {
    Vector v = sub(s.center, r.pos);
    RFloat b = dot(v, r.dir);
    RFloat disc = b * b - dot(v, v) + s.radius * s.radius;
    if (disc < 0.0f) return INF;
    k.disc_nonneg++;
    RFloat d = sqrtf(disc);
    RFloat t2 = b + d;
    if (t2 < 0.0f) return INF;
    RFloat t1 = b - d;
    return t1 > 0.0f ? t1 : t2;
}

// primitive.rs:77-84 Sphere::intersect
inline void sphere_intersect(const Sphere &s, Hit &hit, const Ray &ray, Counters &k) {
    k.leaf_tests++;
    RFloat distance = distance_from_ray(s, ray, k);
    if (distance >= hit.distance) return;
    k.hit_updates++;
    hit.distance = distance;
    hit.pos = normalized(add(ray.pos, sub(mulfed(ray.dir, distance), s.center)));
}

// ---- group.rs:7-20 -------------------------------------------------------
struct Group;
struct Pair {  // Pair::Item / Pair::Group
    bool is_group;
    Sphere item;
    std::unique_ptr<Group> group;
};
struct Group {
    Sphere bound;
    std::vector<Pair> children;
};

// group.rs:28-56 pyramid_recursive
Pair pyramid_recursive(uint32_t level, Vector p, RFloat r) {
    Sphere s{p, r};
    Pair out;
    if (level == 1) {
        out.is_group = false;
        out.item = s;
        return out;
    }
    std::unique_ptr<Group> g(new Group());
    g->children.reserve(5);
    Pair own;
    own.is_group = false;
    own.item = s;
    g->children.push_back(std::move(own));
    g->bound.center = p;
    g->bound.radius = 3.0f * r;

    RFloat rn = 3.0f * r / sqrtf(12.0f);
    const int signs[2] = {-1, 1};
    for (int dz : signs) {
        for (int dx : signs) {
            Vector np = add(p, Vector{(RFloat)dx * rn, rn, (RFloat)dz * rn});
            g->children.push_back(pyramid_recursive(level - 1, np, r * 0.5f));
        }
    }
    out.is_group = true;
    out.item = s;
    out.group = std::move(g);
    return out;
}

// group.rs:72-83 TypedGroup::intersect
void group_intersect(const Group &g, Hit &hit, const Ray &ray, Counters &k) {
    k.bound_tests++;
    if (distance_from_ray(g.bound, ray, k) >= hit.distance) return;
    for (const Pair &c : g.children) {
        if (c.is_group)
            group_intersect(*c.group, hit, ray, k);
        else
            sphere_intersect(c.item, hit, ray, k);
    }
}

// group.rs:95-109 count
void group_count(const Group &g, uint64_t &ng, uint64_t &ni) {
    ng += 1;
    for (const Pair &c : g.children) {
        if (c.is_group)
            group_count(*c.group, ng, ni);
        else
            ni += 1;
    }
}

void group_flatten(const Group &g, std::vector<float> &sph, std::vector<uint32_t> &skip) {
    size_t me = skip.size();
    sph.insert(sph.end(), {g.bound.center.x, g.bound.center.y, g.bound.center.z, g.bound.radius});
    skip.push_back(0);
    for (const Pair &c : g.children) {
        if (c.is_group) {
            group_flatten(*c.group, sph, skip);
        } else {
            sph.insert(sph.end(), {c.item.center.x, c.item.center.y, c.item.center.z, c.item.radius});
            skip.push_back((uint32_t)skip.size() + 1);
        }
    }
    skip[me] = (uint32_t)skip.size();
}

std::unique_ptr<Group> group_from_flat(const float *sph, const uint32_t *skip, uint32_t i) {
    std::unique_ptr<Group> g(new Group());
    g->bound = Sphere{{sph[4 * i], sph[4 * i + 1], sph[4 * i + 2]}, sph[4 * i + 3]};
    uint32_t end = skip[i];
    uint32_t j = i + 1;
    while (j < end) {
        Pair c;
        if (skip[j] > j + 1) {
            c.is_group = true;
            c.item = Sphere{{0, 0, 0}, 0};
            c.group = group_from_flat(sph, skip, j);
        } else {
            c.is_group = false;
            c.item = Sphere{{sph[4 * j], sph[4 * j + 1], sph[4 * j + 2]}, sph[4 * j + 3]};
        }
        g->children.push_back(std::move(c));
        j = skip[j];
    }
    return g;
}

}  // namespace

// ---- render.rs:138-167 Scene ----------------------------------------------
struct orc_scene {
    std::unique_ptr<Group> group;
    Vector directional_light;
    Vector eye;
};

namespace {

enum Kind : uint8_t { K_BACKGROUND = 0, K_AWAY = 1, K_LIT = 2, K_SHADOWED = 3 };

struct TraceStats {
    uint64_t shadow_rays = 0, primary_hits = 0;
};

// render.rs:171-215 Renderer::raytrace
inline RFloat raytrace(const orc_scene &s, const Ray &r, Vector &c, Counters &k, TraceStats &ts, uint8_t *kind) {
    const Vector OBJECT = {(RFloat)0xae / 255.0f, (RFloat)0x31 / 255.0f, (RFloat)0x31 / 255.0f};
    const Vector BACKGROUND = {(RFloat)0x22 / 255.0f, (RFloat)0x0a / 255.0f, (RFloat)0x0a / 255.0f};
    const Vector AMBIENT_OFFSET = {BACKGROUND.x * 0.8f, BACKGROUND.y * 0.8f, BACKGROUND.z * 0.8f};

    Hit h{INF, {0.0f, 0.0f, 0.0f}};
    group_intersect(*s.group, h, r, k);
    if (h.has_missed()) {
        c = add(c, BACKGROUND);
        if (kind) *kind = K_BACKGROUND;
        return 0.0f;
    }
    ts.primary_hits++;
    RFloat g = dot(h.pos, s.directional_light);
    if (g >= 0.0f) {
        c = add(c, AMBIENT_OFFSET);
        if (kind) *kind = K_AWAY;
        return 0.0f;
    }
    // render.rs:199   r.pos + dir*distance + normal*(distance*sqrt(EPSILON))
    const RFloat sqrt_eps = sqrtf(std::numeric_limits<float>::epsilon());
    Vector p = add(add(r.pos, mulfed(r.dir, h.distance)), mulfed(h.pos, h.distance * sqrt_eps));

    ts.shadow_rays++;
    h.distance = INF;  // set_missed
    Ray sr{p, mulfed(s.directional_light, -1.0f)};
    group_intersect(*s.group, h, sr, k);
    if (h.has_missed()) {
        c = add(add(c, mulfed(OBJECT, -g)), AMBIENT_OFFSET);
        if (kind) *kind = K_LIT;
        return 1.0f;
    } else {
        c = add(add(c, BACKGROUND), mulfed(AMBIENT_OFFSET, -g));
        if (kind) *kind = K_SHADOWED;
        return 0.0f;
    }
}

// render.rs:96-103 the `scale` closure: trunc(0.5 + 255 v), >255 -> 255, `as u8` saturates (NaN -> 0)
inline uint8_t scale_u8(RFloat v) {
    RFloat r = 0.5f + 255.0f * v;
    if (r > 255.0f) return 255;
    if (!(r > 0.0f)) return 0;  // negative or NaN: Rust `as u8` saturates to 0
    return (uint8_t)r;
}

// One pixel of render.rs:231-253.
inline void render_pixel(const orc_scene &s, const orc_camera *cam, uint32_t W, uint32_t H, uint32_t spp,
                         uint32_t x, uint32_t y, uint8_t *px, uint8_t *kinds, Counters &k, TraceStats &ts) {
    RFloat ssf = (RFloat)spp;
    RFloat total_recip = recip(ssf * ssf);
    RFloat width = (RFloat)W, height = (RFloat)H;
    Ray ray;
    ray.pos = cam ? Vector{cam->eye[0], cam->eye[1], cam->eye[2]} : s.eye;
    Vector g{0.0f, 0.0f, 0.0f};
    RFloat alpha = 0.0f;
    for (uint32_t ssx = 0; ssx < spp; ssx++) {
        for (uint32_t ssy = 0; ssy < spp; ssy++) {
            RFloat xres = (RFloat)x + (RFloat)ssx / ssf;
            RFloat yres = (RFloat)y + (RFloat)ssy / ssf;
            Vector d;
            d.x = xres - width / 2.0f;
            d.y = (height - yres) - height / 2.0f;
            d.z = width;
            if (cam) {
                // Extension: rotate the camera-space direction by the basis
                // before normalising.  Identity basis gives the reference's numbers.
                Vector w;
                w.x = cam->right[0] * d.x + cam->up[0] * d.y + cam->forward[0] * d.z;
                w.y = cam->right[1] * d.x + cam->up[1] * d.y + cam->forward[1] * d.z;
                w.z = cam->right[2] * d.x + cam->up[2] * d.y + cam->forward[2] * d.z;
                d = w;
            }
            ray.dir = normalized(d);  // Vector::normalize, vec.rs:87-90
            alpha += raytrace(s, ray, g, k, ts, kinds ? kinds + (ssx * spp + ssy) : nullptr);
        }
    }
    g = mulfed(g, total_recip);
    alpha *= total_recip;
    px[0] = scale_u8(g.x);
    px[1] = scale_u8(g.y);
    px[2] = scale_u8(g.z);
    px[3] = scale_u8(alpha);
}

void accumulate(orc_counters *out, const Counters &k, const TraceStats &ts, uint64_t primary) {
    if (!out) return;
    out->primary_rays += primary;
    out->shadow_rays += ts.shadow_rays;
    out->primary_hits += ts.primary_hits;
    out->bound_tests += k.bound_tests;
    out->leaf_tests += k.leaf_tests;
    out->disc_nonneg += k.disc_nonneg;
    out->hit_updates += k.hit_updates;
}

}  // namespace

extern "C" {

orc_scene *orc_scene_create(uint32_t level, const float origin[3], float radius,
                            const float light_unnormalised[3], const float eye[3]) {
    if (level <= 1) return nullptr;  // group.rs:59-60 assert!(level > 1)
    orc_scene *s = new orc_scene();
    Pair root = pyramid_recursive(level, Vector{origin[0], origin[1], origin[2]}, radius);
    s->group = std::move(root.group);
    s->directional_light = normalized(Vector{light_unnormalised[0], light_unnormalised[1], light_unnormalised[2]});
    s->eye = Vector{eye[0], eye[1], eye[2]};
    return s;
}

orc_scene *orc_scene_create_default(void) {
    const float origin[3] = {0.0f, -1.0f, 0.0f};
    const float light[3] = {-1.0f, -3.0f, 2.0f};
    const float eye[3] = {0.0f, 0.0f, -4.0f};
    return orc_scene_create(8, origin, 1.0f, light, eye);
}

orc_scene *orc_scene_create_from_nodes(uint32_t n, const float *spheres4, const uint32_t *skip,
                                       const float light[3], const float eye[3]) {
    if (n < 2 || skip[0] != n) return nullptr;
    orc_scene *s = new orc_scene();
    s->group = group_from_flat(spheres4, skip, 0);
    s->directional_light = Vector{light[0], light[1], light[2]};
    s->eye = Vector{eye[0], eye[1], eye[2]};
    return s;
}

void orc_scene_destroy(orc_scene *s) { delete s; }

void orc_scene_counts(const orc_scene *s, uint64_t *groups, uint64_t *items) {
    uint64_t ng = 0, ni = 0;
    group_count(*s->group, ng, ni);
    *groups = ng;
    *items = ni;
}

uint32_t orc_scene_flatten(const orc_scene *s, float *spheres4, uint32_t *skip, uint32_t cap) {
    std::vector<float> sph;
    std::vector<uint32_t> sk;
    group_flatten(*s->group, sph, sk);
    uint32_t n = (uint32_t)sk.size();
    uint32_t m = n < cap ? n : cap;
    if (spheres4) memcpy(spheres4, sph.data(), sizeof(float) * 4 * m);
    if (skip) memcpy(skip, sk.data(), sizeof(uint32_t) * m);
    return n;
}

void orc_scene_light(const orc_scene *s, float light[3]) {
    light[0] = s->directional_light.x;
    light[1] = s->directional_light.y;
    light[2] = s->directional_light.z;
}

void orc_scene_eye(const orc_scene *s, float eye[3]) {
    eye[0] = s->eye.x;
    eye[1] = s->eye.y;
    eye[2] = s->eye.z;
}

float orc_sphere_distance_from_ray(const float center[3], float radius, const orc_ray *r) {
    Counters k;
    Sphere s{{center[0], center[1], center[2]}, radius};
    Ray ray{{r->pos[0], r->pos[1], r->pos[2]}, {r->dir[0], r->dir[1], r->dir[2]}};
    return distance_from_ray(s, ray, k);
}

void orc_sphere_intersect(const float center[3], float radius, orc_hit *h, const orc_ray *r) {
    Counters k;
    Sphere s{{center[0], center[1], center[2]}, radius};
    Ray ray{{r->pos[0], r->pos[1], r->pos[2]}, {r->dir[0], r->dir[1], r->dir[2]}};
    Hit hit{h->distance, {h->normal[0], h->normal[1], h->normal[2]}};
    sphere_intersect(s, hit, ray, k);
    h->distance = hit.distance;
    h->normal[0] = hit.pos.x;
    h->normal[1] = hit.pos.y;
    h->normal[2] = hit.pos.z;
}

void orc_vec_normalized(const float v[3], float out[3]) {
    Vector n = normalized(Vector{v[0], v[1], v[2]});
    out[0] = n.x;
    out[1] = n.y;
    out[2] = n.z;
}

float orc_vec_len(const float v[3]) { return len(Vector{v[0], v[1], v[2]}); }

void orc_trace_rays(const orc_scene *s, size_t n, const orc_ray *rays, orc_hit *hits) {
    Counters k;
    for (size_t i = 0; i < n; i++) {
        Ray ray{{rays[i].pos[0], rays[i].pos[1], rays[i].pos[2]}, {rays[i].dir[0], rays[i].dir[1], rays[i].dir[2]}};
        Hit h{INF, {0.0f, 0.0f, 0.0f}};
        group_intersect(*s->group, h, ray, k);
        hits[i].distance = h.distance;
        hits[i].normal[0] = h.pos.x;
        hits[i].normal[1] = h.pos.y;
        hits[i].normal[2] = h.pos.z;
    }
}

void orc_render_region(const orc_scene *s, const orc_camera *cam, uint32_t width, uint32_t height, uint32_t spp,
                       uint32_t l, uint32_t b, uint32_t r, uint32_t t, uint8_t *rgba_out, uint8_t *kinds_out,
                       orc_counters *ctr) {
    Counters k;
    TraceStats ts;
    uint32_t rw = r - l;
    uint32_t nsamp = spp * spp;
    for (uint32_t y = b; y < t; y++) {
        for (uint32_t x = l; x < r; x++) {
            size_t ofs = (size_t)(y - b) * rw + (x - l);  // render.rs:69-71 buffer_offset
            render_pixel(*s, cam, width, height, spp, x, y, rgba_out + ofs * 4,
                         kinds_out ? kinds_out + ofs * nsamp : nullptr, k, ts);
        }
    }
    accumulate(ctr, k, ts, (uint64_t)rw * (t - b) * nsamp);
}

void orc_render_rows(const orc_scene *s, const orc_camera *cam, uint32_t width, uint32_t height, uint32_t spp,
                     uint32_t row_start, uint32_t row_stride, uint32_t row_count, uint32_t nthreads,
                     uint8_t *rgba_out, orc_counters *ctr) {
    const uint32_t CHUNK = 64;  // render.rs:264
    uint32_t bx = (width + CHUNK - 1) / CHUNK;
    uint32_t by = (row_count + CHUNK - 1) / CHUNK;
    uint32_t nb = bx * by;
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint32_t> next(0);
    std::vector<Counters> ks(nthreads);
    std::vector<TraceStats> tss(nthreads);
    auto worker = [&](uint32_t tid) {
        for (;;) {
            uint32_t i = next.fetch_add(1);  // FIFO bucket queue, row-major y then x (render.rs:273-298)
            if (i >= nb) break;
            uint32_t x0 = (i % bx) * CHUNK, j0 = (i / bx) * CHUNK;
            uint32_t x1 = x0 + CHUNK < width ? x0 + CHUNK : width;
            uint32_t j1 = j0 + CHUNK < row_count ? j0 + CHUNK : row_count;
            for (uint32_t j = j0; j < j1; j++) {
                uint32_t y = row_start + j * row_stride;
                for (uint32_t x = x0; x < x1; x++)
                    render_pixel(*s, cam, width, height, spp, x, y, rgba_out + ((size_t)j * width + x) * 4, nullptr,
                                 ks[tid], tss[tid]);
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t i = 1; i < nthreads; i++) th.emplace_back(worker, i);
    worker(0);
    for (auto &t : th) t.join();
    for (uint32_t i = 0; i < nthreads; i++)
        accumulate(ctr, ks[i], tss[i], 0);
    if (ctr) ctr->primary_rays += (uint64_t)width * row_count * spp * spp;
}

void orc_render(const orc_scene *s, const orc_camera *cam, uint32_t width, uint32_t height, uint32_t spp,
                uint32_t nthreads, uint8_t *rgba_out, orc_counters *ctr) {
    orc_render_rows(s, cam, width, height, spp, 0, 1, height, nthreads, rgba_out, ctr);
}

}  // extern "C"
