// rt_scene.h -- host-side scene flattening (replaces group.rs:28-66 + render.rs:145-166).
#pragma once
#include <stdint.h>
#include <vector>

namespace rt {

struct FlatScene {
    std::vector<float> sph;      // n x {cx, cy, cz, r}: pre-order; a group contributes its bound first
    std::vector<uint32_t> skip;  // group: first index after its subtree; leaf: i + 1
    uint64_t groups = 0, items = 0;
    float light[3] = {0, 0, 0};  // normalised
    float eye[3] = {0, 0, 0};
    uint32_t level = 0;  // 0 for hand-built trees
    float leaf_rmin = 0.0f;  // smallest leaf radius
};

// Nodes in a pyramid subtree of the given level: S(1) = 1, S(L) = 2 + 4 S(L-1).
uint64_t pyramid_subtree_nodes(uint32_t level);

// pyramid(level, origin, radius) flattened in the reference's child order
// [own sphere, (dz-,dx-), (dz-,dx+), (dz+,dx-), (dz+,dx+)] (group.rs:39,44-52).
void flatten_pyramid(uint32_t level, const float origin[3], float radius, FlatScene &out);

// Validates a caller-supplied pre-order tree; returns false with a reason on malformed input.
bool validate_nodes(uint32_t n, const uint32_t *skip, uint64_t *groups, uint64_t *items, const char **why);

// Vector::normalized (vec.rs:93-95) on the host, in strict f32.
void normalize3(const float v[3], float out[3]);

}  // namespace rt
