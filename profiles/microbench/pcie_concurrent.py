"""Raw concurrent pinned-memory copy bandwidth of this box: every rank copies `MB` megabytes host->device and
device->host at the same time (two streams), all ranks together.  The ceiling the end-to-end step (560 MB of host
traffic per rank and step) can reach at 1 / 2 / 4 / 8 ranks.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/microbench/pcie_concurrent.py"""
import json, os, sys
import torch
import torch.distributed as td

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 280
n = MB * 1024 * 1024 // 4
h_in, h_out = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
d_in, d_out = torch.empty(n, device="cuda"), torch.ones(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("h2d", "d2h", "both"):
    for it in range(3):
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    byts = MB * 1024 * 1024 * (2 if mode == "both" else 1)
    res[mode] = {"ms_max_over_ranks": float(t.item()), "gbs_per_rank": byts / (float(t.item()) * 1e-3) / 1e9,
                 "gbs_aggregate": world * byts / (float(t.item()) * 1e-3) / 1e9}
if rank == 0:
    print(json.dumps({"ranks": world, "mb_per_direction": MB, **res}))
if world > 1:
    td.destroy_process_group()
