#!/usr/bin/env python
"""bench.py -- throughput of the rust-tracer hot path on B200 (one JSON line).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A *step* is one pass of the hot path over one frame: ray generation, primary and
shadow traversal, shading, supersample averaging and RGBA8 quantisation of every
pixel.  Workload at N=1 (BASELINE.json configs[1], "C2"): the reference's ~20k-sphere
scene (level 8, 21,845 spheres) at 3840x2160, 1 sample per pixel.  With N>1 every
rank renders whole frames of that workload (frame-sharded sweep, BASELINE C5's
partition): no data-path collective, weak scaling.  `--mode bands` instead splits ONE
frame into interleaved row bands gathered to rank 0 over NCCL (BASELINE C4's
partition, strong scaling).  `--workload c5` is BASELINE configs[4] itself: the 120-frame
orbit of the camera about the flake over the level-9 scene at 3840x2160, 4x4 samples,
step i of rank r rendering frame (i*N + r) mod 120 (c1/c3/c4 select the other configs).

metric  = Mrays/s (primary + shadow rays, counted as the reference's work is counted)
value   = whole-job rays / device time of the K steps (CUDA events on the launch
          stream, max over ranks), frames written to HBM, nothing leaves the GPU
e2e     = the same metric through the C ABI with a HOST output buffer
          (rt_render_frame): kernel-parameter block in, frame out over PCIe, every step
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))

WORKLOADS = {
    # name: (width, height, spp, level)
    "c1": (1024, 768, 4, 8),     # `make image`
    "c2": (3840, 2160, 1, 8),    # BASELINE configs[1]: the headline single-GPU 4K render
    "c3": (3840, 2160, 4, 9),    # deeper flake (87,381 spheres), 4x4 supersampling
    "c3l10": (3840, 2160, 4, 10),
    "c4": (7680, 4320, 4, 9),
    "c5": (3840, 2160, 4, 9),    # BASELINE configs[4]: 120-frame orbit sweep of C3-sized frames, frame f -> rank f mod N
}
ORBIT_FRAMES = 120               # c5: eye and camera basis rotated about the flake's axis by 2 pi f / 120
# algorithmic flop per ray of REFERENCE work (17 T + 3 P + 19 U + 20 f_p + 30 f_s),
# from the oracle's counters (tests/golden/oracle_derived.json); used when the fixture lacks the config.
FLOP_PER_RAY_FALLBACK = 645.0
METRIC = "Mrays/s (primary+shadow)"


def orbit_basis(frame, n_frames=ORBIT_FRAMES, eye=(0.0, 0.0, -4.0)):
    """(eye, right, up, forward) of orbit frame `frame`: the reference camera (render.rs:145-166, 238-243)
    rotated about the vertical axis; frame 0 is the reference camera exactly.  Same arithmetic as
    rtrace_b200.orbit_camera and the CLI's --frames (host/main.cpp)."""
    import math
    th = 2.0 * math.pi * frame / n_frames
    c, s = (1.0, 0.0) if frame % n_frames == 0 else (math.cos(th), math.sin(th))

    def rot(v):
        return (c * v[0] + s * v[2], v[1], -s * v[0] + c * v[2])
    return rot(eye), rot((1, 0, 0)), (0, 1, 0), rot((0, 0, 1))


def workload_name(name, level, width, height, spp):
    s = "%s: pyramid level %d (%d spheres) at %dx%d, %d spp" % (name, level, (4 ** level - 1) // 3, width, height,
                                                               spp * spp)
    if name == "c5":
        s += ", %d-frame orbit sweep (step i on rank r = frame (i*N + r) mod %d)" % (ORBIT_FRAMES, ORBIT_FRAMES)
    return s


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def flop_per_ray(width, height, spp, level):
    try:
        cases = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_derived.json")))["cases"]
        for c in cases:
            if (c["width"], c["height"], c["spp"], c["level"]) == (width, height, spp, level):
                return c["flop_per_ray"], "oracle counters for this exact config (tests/golden/oracle_derived.json)"
        for c in cases:  # same resolution, other level: work per ray is flat in depth
            if (c["width"], c["height"]) == (width, height):
                return c["flop_per_ray"], "oracle counters at %dx%d level %d" % (width, height, c["level"])
    except Exception:
        pass
    return FLOP_PER_RAY_FALLBACK, "SURVEY 8(d) figure for C2"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the benchmark runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin, t_end):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t_begin - 0.05 <= t <= t_end + 0.15 and len(r) >= 7] or \
               [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]

        def num(s):
            try:
                return float(s)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        pw = [num(r[2]) for r in rows if num(r[2]) is not None]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": num(rows[0][1]),
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


class stdout_to_stderr:
    """File descriptor 1 points at stderr inside the block.  NCCL prints "NCCL version ..." on stdout when the
    first communicator is created; bench.py's stdout carries exactly one JSON line, so that goes to stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def pin_to_gpu_cpus(local_rank):
    """Multi-GPU runs: keep this rank's host threads (and so its pinned frame buffers, first touch) on the
    CPUs NVML reports as local to its GPU, so that eight ranks copying frames out do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to GPU %d" % (len(cpus), idx)
    except Exception as e:   # best effort
        return "not pinned (%s)" % type(e).__name__
    return "not pinned"


def cpu_leg(width, height, spp, level, min_seconds, row_stride):
    """The oracle (CPU restatement of the reference algorithm) on every host core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as o   # bench.py's cpu_baseline / reference legs are allowed to run the oracle
    cores = os.cpu_count() or 1
    s = o.Scene(level=level)
    rows = (height + row_stride - 1) // row_stride
    rays = secs = 0.0
    reps = 0
    ctr = None
    while secs < min_seconds or reps == 0:
        t0 = time.perf_counter()
        _, ctr = s.render_rows(width, height, spp, 0, row_stride, rows, threads=cores)
        secs += time.perf_counter() - t0
        rays += ctr.primary_rays + ctr.shadow_rays
        reps += 1
    sample = "rows 0,%d,%d,.. (%d of %d rows) of the %dx%d spp %d level %d frame, %d repeat(s), %.1f s" % (
        row_stride, 2 * row_stride, rows, height, width, height, spp, level, reps, secs)
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample,
            "flop_per_ray": ctr.flop_per_ray()}, rays / reps, secs / reps


def run_reference(args, width, height, spp, level):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust binary cannot be built
    here: no cargo/rustc) on all host threads, same metric and config.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as o
    cores = os.cpu_count() or 1
    s = o.Scene(level=level)
    stride = 4   # each step renders every 4th row: an unbiased quarter of the frame
    rows = (height + stride - 1) // stride
    for _ in range(min(args.warmup, 1)):
        s.render_rows(width, height, spp, 0, stride, rows, threads=cores)
    rays = 0
    t0 = time.perf_counter()
    budget_steps = args.steps
    done = 0
    for i in range(budget_steps):
        cam = None
        if args.workload == "c5":   # orbit sweep: step i is frame i of the orbit (same cameras as the native arm)
            cam = o.make_camera(*orbit_basis(i % ORBIT_FRAMES))
        _, ctr = s.render_rows(width, height, spp, 0, stride, rows, threads=cores, camera=cam)
        rays += ctr.primary_rays + ctr.shadow_rays
        done += 1
        if time.perf_counter() - t0 > 150.0:   # keep the whole run within a few minutes
            break
    secs = time.perf_counter() - t0
    v = rays / secs / 1e6
    sample = "each step = rows 0,4,8,.. (%d of %d) of the frame on %d host threads; %d steps timed" % (rows, height, cores, done)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 1), "ms_per_step": secs / done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, level, width, height, spp)},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="frames", choices=["frames", "bands"],
                    help="N>1: 'frames' = each rank renders whole frames (weak); 'bands' = one frame split "
                         "into interleaved rows gathered to rank 0 over NCCL (strong)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="bands mode: 'peer' = every rank's kernel stores its rows straight into rank 0's frame "
                         "through CUDA-IPC peer memory over NVLink; 'nccl' = dist.gather of the bands + de-interleave")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-l2-flush", action="store_true")
    args = ap.parse_args()
    width, height, spp, level = WORKLOADS[args.workload]
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3   # timing rule: at least 3 warm-up steps

    if args.impl == "reference":
        run_reference(args, width, height, spp, level)
        return

    import torch
    import torch.distributed as dist
    import rtrace_b200 as rt   # raises ImportError if the CUDA library is not built: no fallback

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_note = pin_to_gpu_cpus(local_rank) if world > 1 and not os.environ.get("RTRACE_NO_PIN") else None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    rt.set_device(local_rank)
    if world > 1:
        with stdout_to_stderr():   # communicator creation and the first collective: NCCL's banner
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
    rt.set_variant(args.variant)

    scene = rt.Scene(level=level)
    opts = rt.RenderOptions(width, height, spp)
    stream = torch.cuda.current_stream()
    bands = args.mode == "bands" and world > 1
    sweep = args.workload == "c5"
    if sweep and bands:
        raise SystemExit("bench.py: c5 is the frame-sharded orbit sweep; use --mode frames")
    cams = [rt.make_camera(*orbit_basis(f)) for f in range(ORBIT_FRAMES)] if sweep else None

    def frame_of(i):   # c5: step i of this rank renders orbit frame (i * world + rank) mod 120
        return (i * world + rank) % ORBIT_FRAMES
    from rtrace_b200 import partition
    if bands:
        row_start, row_stride, my_rows = partition.band_spec(height, rank, world)
        max_rows = partition.band_capacity(height, world)
    else:
        my_rows, max_rows, row_start, row_stride = height, height, 0, 1
    fb = torch.zeros((max_rows + 16, width, 4), dtype=torch.uint8, device="cuda")   # + one row block (blocked partition)
    gathered = [torch.zeros_like(fb[:max_rows]) for _ in range(world)] if (bands and rank == 0) else None
    frame_box = [None]
    peer = bands and args.gather == "peer"
    peer_ptr, peer_pitch = None, 0
    if peer:
        # rank 0 owns the frame; the others map it through a CUDA IPC handle and write their rows into it
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            frame_base = rt.device_alloc(height * width * 4)
            handle.copy_(torch.frombuffer(bytearray(rt.ipc_export(frame_base)), dtype=torch.uint8))
        dist.broadcast(handle, src=0)
        if rank != 0:
            frame_base = rt.ipc_open(bytes(handle.cpu().numpy().tobytes()))
        # blocks of 16 consecutive rows per rank (whole cull tiles stay contiguous in the image)
        row_start, row_stride, row_block, my_rows = partition.block_band_spec(height, rank, world, 16)

    # kernels launched per step (the PHASED variant is four launches per frame)
    _, st0 = rt.Renderer.render_rows(opts, scene, row_start=0, row_stride=1, row_count=min(my_rows, height),
                                     out_ptr=fb.data_ptr(), stream=stream.cuda_stream, want_stats=True)
    launches_per_step = int(st0.kernel_launches)
    # rays of this rank's share of a step, counted on the device the way the reference's work is counted
    if peer:   # blocked partition: count whole-frame rays once; each rank gets its share by pixel count
        primary_all, shadow_all = scene.count_rays(width, height, spp)
        primary = width * my_rows * spp * spp
        shadow = int(round(shadow_all * my_rows / height))
    else:
        primary, shadow = scene.count_rays(width, height, spp, row_start, row_stride, my_rows)
    rays_rank = primary + shadow
    e2e_steps = max(5, min(args.steps, 100))
    e2e_rays_rank = rays_rank
    if sweep:   # every orbit frame casts its own number of shadow rays: count each frame this rank renders
        table = {}
        for f in sorted({frame_of(i) for i in range(max(args.steps, e2e_steps))}):
            table[f] = scene.count_rays(width, height, spp, camera=cams[f])
        primary = sum(table[frame_of(i)][0] for i in range(args.steps)) / args.steps
        shadow = sum(table[frame_of(i)][1] for i in range(args.steps)) / args.steps
        rays_rank = primary + shadow                      # mean per timed step
        e2e_rays_rank = sum(sum(table[frame_of(i)]) for i in range(e2e_steps)) / e2e_steps
    flush = None if args.no_l2_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(i=0):
        if sweep:
            rt.Renderer.render_rows(opts, scene, camera=cams[frame_of(i)], out_ptr=fb.data_ptr(),
                                    stream=stream.cuda_stream)
            return
        if peer:   # the traversal kernels' framebuffer stores ARE the gather (NVLink peer writes)
            rt.Renderer.render_row_blocks(opts, scene, row_start, row_stride, row_block, my_rows, frame_base,
                                          pitch=width * 4, absolute_rows=True, stream=stream.cuda_stream)
            return
        rt.Renderer.render_rows(opts, scene, row_start=row_start, row_stride=row_stride, row_count=my_rows,
                                out_ptr=fb.data_ptr(), stream=stream.cuda_stream)
        if bands:
            dist.gather(fb[:max_rows], gathered, dst=0)
            if rank == 0:   # de-interleave: row r*world + g  <-  band g row r
                frame_box[0] = partition.deinterleave(gathered, height)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]) if rank == 0 else None
    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- value: K steps, device time from CUDA events on the launch stream --------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin = time.time()
    w0 = time.perf_counter()
    for i, (a, b) in enumerate(ev):
        if flush is not None:
            flush.zero_()          # evict L2 (126 MB) between timed steps; outside the event pair
        a.record(stream)
        step(i)
        b.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    peer_ok = None
    if peer:
        barrier()
        if rank == 0:   # the gathered frame must equal a frame rendered by rank 0 alone
            import numpy as _np
            whole = torch.zeros((height, width, 4), dtype=torch.uint8, device="cuda")
            rt.Renderer.render_rows(opts, scene, out_ptr=whole.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            got = rt.PinnedBuffer(height * width * 4)
            rt.memcpy(got.ptr, frame_base, got.nbytes)
            peer_ok = bool(_np.array_equal(got.array.reshape(height, width, 4), whole.cpu().numpy()))
        barrier()

    # ---- e2e: the C-ABI call a user makes, every frame delivered to HOST memory ----------------
    # frames mode: rt_render_sweep (what `rtrace --frames` calls): the device-to-host copy of frame f
    # overlaps the render of frame f+1; the callback sees every frame in pinned host memory.
    # bands mode: rt_render_rows into a pinned host buffer, synchronously.
    seen = []

    def on_frame(f, arr):
        seen.append(int(arr[0, 0, 0]))   # touch the host copy of every frame

    e_start, e_stride, e_rows = partition.band_spec(height, rank, world) if bands else (0, 1, height)
    pinned = rt.PinnedBuffer(max(e_rows, 1) * width * 4) if bands else None

    def e2e_run(n):
        if bands:
            for _ in range(n):
                rt.Renderer.render_rows(opts, scene, row_start=e_start, row_stride=e_stride, row_count=e_rows,
                                        out_ptr=pinned.ptr)
        else:
            rt.Renderer.render_sweep(opts, scene, n, on_frame=on_frame, rgb=True,
                                     cameras=[cams[frame_of(i)] for i in range(n)] if sweep else None)

    e2e_run(3)
    barrier()
    e0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if sampler else None

    # ---- max over ranks ------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_ms, float(rays_rank), float(e2e_rays_rank)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms, rays_job, e2e_rays_job = tmax[0].item(), tmax[1].item(), tsum[2].item(), tsum[3].item()
    else:
        rays_job, e2e_rays_job = float(rays_rank), float(e2e_rays_rank)

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = rays_job / (ms_per_step * 1e-3) / 1e6
        e2e_value = e2e_rays_job / (e2e_ms / e2e_steps * 1e-3) / 1e6
        fpr, fpr_src = flop_per_ray(width, height, spp, level)
        try:
            peak_tf, eff_mhz = rt.measure_fp32_peak(local_rank)
            peak_src = "measured live: FFMA chains on all SMs (rt_measure_fp32_peak), effective %.0f MHz" % eff_mhz
        except Exception as e:   # pragma: no cover
            peak_tf, peak_src = 148 * 128 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 128 lanes x 2 x 1965 MHz (%s)" % e
        # per-launch figures for the dominant (only) kernel: one launch per step per rank
        rays_launch = rays_job / world if not bands else rays_job / world
        achieved_tf = rays_launch * fpr / (ms_per_step * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        fb_bytes = my_rows * width * 4
        traffic, ncu_facts = None, None
        try:
            summ = json.load(open(os.path.join(ROOT, "profiles", "latest_summary.json"))).get(args.workload, {})
            traffic = summ.get("dram_bytes_per_launch")
            ncu_facts = {k: summ[k] for k in ("dominant_kernel", "fma_pipe_cycles_active_pct", "issue_active_pct",
                                              "warp_instructions_per_frame", "source") if k in summ} or None
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if bands else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args.workload, level, width, height, spp),
                "partition": (("interleaved row bands; kernels store into rank 0's frame through IPC peer memory (NVLink)"
                               " in blocks of 16 rows" if peer else "interleaved row bands + NCCL gather to rank 0") if bands else
                              "one whole frame per rank per step (frame-sharded sweep), no collective"),
                "l2": "not flushed" if flush is None else "flushed between steps (256 MiB memset, outside the timed events)",
                "rays_per_frame": {"primary": primary * (world if bands else 1), "shadow": None if bands else shadow},
                "mpixels_per_s": width * height * (1 if bands else world) / (ms_per_step * 1e-3) / 1e6,
                "variant": args.variant, "wall_ms_per_step_incl_flush": wall_ms / args.steps, "host_cpus": cpu_note,
            },
            "roofline": {
                "bound": "fp32", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "ncu": ncu_facts,   # what the hardware actually did (committed ncu capture), next to the reference-work figure
                "note": "path is FP32-issue bound, not HBM/tensor (SURVEY 8d): achieved = rays/launch x %.1f "
                        "algorithmic flop/ray of REFERENCE work (%s) / event time; peak = %s" % (fpr, fpr_src, peak_src),
                "hbm": {"achieved": fb_bytes / (ms_per_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": fb_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak,
                        "note": "algorithmic HBM bytes = framebuffer written once (4 B/pixel); peak = MEASURED_PEAKS.json hbm_gbs" if peaks else "peak = fallback 6.65 TB/s"},
            },
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 512,
                    "d2h_bytes_per_step": fb_bytes if bands else fb_bytes // 4 * 3, "steps": e2e_steps,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "note": ("rt_render_rows -> pinned host buffer, synchronous" if bands else
                             "rt_render_sweep_rgb (what `rtrace --frames` calls): per frame the kernel-parameter blocks "
                             "(camera, options) go in and the RGB8 frame -- the body of the reference's P6 file, its sink "
                             "drops alpha (render.rs:389-397) -- comes out to pinned host memory; copy of frame f "
                             "overlaps the render of frame f+1")},
            "gpu_launches": args.steps * world * launches_per_step,
            "gathered_frame_verified": peer_ok,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cb, _, _ = cpu_leg(width, height, spp, level, 10.0, 1)
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
