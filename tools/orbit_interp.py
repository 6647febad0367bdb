"""Per-orbit-frame counters from a strided capture (tools/round_profiles.sh with C5_STRIDE > 1).

`captured[i]` belongs to orbit frame i * stride; frames in between are interpolated linearly over the orbit angle,
and the orbit closes: the frame after the last one is frame 0 again."""


def stride_of(n_frames, n_captured):
    """Smallest stride whose capture of an n_frames orbit has n_captured frames (None if there is none)."""
    return next((st for st in range(1, n_frames + 1) if len(range(0, n_frames, st)) == n_captured), None)


def interpolate_orbit(captured, n_frames, stride):
    out = []
    n_cap = len(captured)
    for f in range(n_frames):
        i, r = divmod(f, stride)
        a = captured[i]
        b = captured[i + 1] if (i + 1) * stride < n_frames and i + 1 < n_cap else captured[0]
        span = min(stride, n_frames - i * stride)
        out.append(a + (b - a) * r / span)
    return out
