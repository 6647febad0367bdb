"""Copy-only D2H probe under torchrun: every rank copies at once (ring of 3 buffers, 200 copies each);
prints every rank's own rate, so that the spread between the host links under contention is visible."""
import os, sys, json, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rust-tracer_b200"))
import torch, torch.distributed as dist
import rtrace_b200 as rt
r, w, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); rt.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for nbytes, iters in ((24883200, 200), (24883200, 200), (12441600, 400)):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = rt.microbench_d2h(nbytes, iters, 3)
    el = time.perf_counter() - t0
    t = torch.zeros(w, dtype=torch.float64, device="cuda"); t[r] = g
    e = torch.zeros(w, dtype=torch.float64, device="cuda"); e[r] = el
    dist.all_reduce(t); dist.all_reduce(e)
    if r == 0:
        rates = [round(x, 1) for x in t.tolist()]
        print(json.dumps({"world": w, "bytes": nbytes, "iters": iters, "rank_gbs": rates, "sum": round(sum(rates), 1),
                          "world_x_min": round(min(rates) * w, 1), "true_aggregate": round(nbytes * iters * w / max(e.tolist()) / 1e9, 1)}), flush=True)
dist.destroy_process_group()
