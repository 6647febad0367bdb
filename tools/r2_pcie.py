"""Copy-only D2H probe under torchrun: every rank copies at once; ring sizes 1..4, two frame sizes."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rust-tracer_b200"))
import torch, torch.distributed as dist
import rtrace_b200 as rt
r, w, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); rt.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for nbytes in (24883200, 12441600):
    for nb in (1, 2, 3, 4, 6):
        dist.barrier(); torch.cuda.synchronize()
        g = rt.microbench_d2h(nbytes, 48, nb)
        t = torch.tensor([g], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        if r == 0:
            print(json.dumps({"world": w, "bytes": nbytes, "ring": nb, "aggregate_gbs": round(t.item(), 1), "rank0_gbs": round(g, 1)}), flush=True)
dist.destroy_process_group()
