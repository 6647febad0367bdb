// rt_scene.cpp -- builds the flattened sphere hierarchy on the host.
//
// The reference builds a heap tree of Vec<Pair> (group.rs:28-56) and recurses
// over it per ray.  Here the same spheres are emitted straight into a pre-order
// array with skip links, which is what the stackless GPU walk consumes.  The
// pyramid is perfectly regular, so every skip link is known from the level alone
// (S(L) = 2 + 4 S(L-1)) and no tree is ever materialised.
//
// Compiled with -ffp-contract=off: centres and radii must carry the reference's
// f32 roundings (rn = 3 r / sqrt(12), child radius r/2, bound radius 3 r).
#include "rt_scene.h"

#include <cmath>

namespace rt {

uint64_t pyramid_subtree_nodes(uint32_t level) {
    uint64_t s = 1;
    for (uint32_t l = 2; l <= level; l++) s = 2 + 4 * s;
    return s;
}

namespace {
struct Frame {
    float x, y, z, r;
    uint32_t level;
};
}  // namespace

void flatten_pyramid(uint32_t level, const float origin[3], float radius, FlatScene &out) {
    const uint64_t n = pyramid_subtree_nodes(level);
    out.sph.clear();
    out.skip.clear();
    out.sph.reserve(n * 4);
    out.skip.reserve(n);
    out.groups = out.items = 0;
    out.level = level;
    out.leaf_rmin = radius;
    for (uint32_t l = 1; l < level; l++) out.leaf_rmin *= 0.5f;

    // Explicit DFS stack; children are pushed in reverse so they pop in reference order.
    std::vector<Frame> stack;
    stack.push_back(Frame{origin[0], origin[1], origin[2], radius, level});
    const float sqrt12 = sqrtf(12.0f);
    while (!stack.empty()) {
        Frame f = stack.back();
        stack.pop_back();
        uint32_t i = (uint32_t)out.skip.size();
        if (f.level == 1) {  // group.rs:33-35: a bare sphere
            out.sph.insert(out.sph.end(), {f.x, f.y, f.z, f.r});
            out.skip.push_back(i + 1);
            out.items++;
            continue;
        }
        // group.rs:40-41: the bound (p, 3r), then the group's own sphere as child 0 (group.rs:39)
        out.sph.insert(out.sph.end(), {f.x, f.y, f.z, 3.0f * f.r});
        out.skip.push_back(i + (uint32_t)pyramid_subtree_nodes(f.level));
        out.groups++;
        out.sph.insert(out.sph.end(), {f.x, f.y, f.z, f.r});
        out.skip.push_back(i + 2);
        out.items++;
        const float rn = 3.0f * f.r / sqrt12;  // group.rs:43
        const float half = f.r * 0.5f;         // group.rs:52
        for (int dz = 1; dz >= -1; dz -= 2)    // reversed: popped as dz = -1, +1 / dx = -1, +1
            for (int dx = 1; dx >= -1; dx -= 2)
                stack.push_back(Frame{f.x + (float)dx * rn, f.y + rn, f.z + (float)dz * rn, half, f.level - 1});
    }
}

bool validate_nodes(uint32_t n, const uint32_t *skip, uint64_t *groups, uint64_t *items, const char **why) {
    *groups = *items = 0;
    if (n < 2) {
        *why = "a scene needs a root group and at least one item";
        return false;
    }
    if (skip[0] != n) {
        *why = "node 0 must be a group spanning all n nodes (skip[0] == n)";
        return false;
    }
    // Every skip link must point forward, inside its parent's span.
    std::vector<uint32_t> ends;
    for (uint32_t i = 0; i < n; i++) {
        while (!ends.empty() && ends.back() == i) ends.pop_back();
        uint32_t limit = ends.empty() ? n : ends.back();
        if (skip[i] <= i || skip[i] > limit) {
            *why = "skip link out of range or crossing its parent's subtree";
            return false;
        }
        if (skip[i] > i + 1) {
            (*groups)++;
            ends.push_back(skip[i]);
        } else {
            (*items)++;
        }
    }
    return true;
}

void normalize3(const float v[3], float out[3]) {
    float d = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float inv = 1.0f / sqrtf(d);
    out[0] = v[0] * inv;
    out[1] = v[1] * inv;
    out[2] = v[2] * inv;
}

}  // namespace rt
