#!/usr/bin/env python3
"""The host-fabric ceiling behind the end-to-end numbers: plain pinned-memory copies, no kernels.
Every rank copies BYTES host -> device and device -> host at the same time on two streams (what the
e2e leg of bench.py does around the front-end), all ranks together; prints per-rank and total GB/s.

    python tools/pcie_ceiling.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
           --master-port 29555 tools/pcie_ceiling.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                  # noqa: E402
import torch.distributed as dist              # noqa: E402
import bench                                  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = bench.pin_to_gpu_numa_node(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

H2D, D2H = 37_627_560, 26_107_440              # bytes per step of config #2's e2e leg
h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in = torch.empty(H2D, dtype=torch.uint8, device=dev)
d_out = torch.empty(D2H, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, do_in, do_out in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
    for _ in range(3):
        if do_in:
            d_in.copy_(h_in, non_blocking=True)
        if do_out:
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 40
    e0.record()
    s1.wait_event(e0)
    s2.wait_event(e0)
    for _ in range(iters):
        if do_in:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    res[name] = {"h2d_gbs_per_rank": (H2D * iters / (ms / 1e3) / 1e9) if do_in else 0.0,
                 "d2h_gbs_per_rank": (D2H * iters / (ms / 1e3) / 1e9) if do_out else 0.0,
                 "steps_per_s_per_rank": iters / (ms / 1e3)}
if rank == 0:
    both = res["both"]
    print(json.dumps({"n_gpus": world, "numa_local_cpus": numa, "bytes_per_step": {"h2d": H2D, "d2h": D2H}, "copies": res,
                      "total_gbs_both_directions": world * (both["h2d_gbs_per_rank"] + both["d2h_gbs_per_rank"]),
                      "e2e_ceiling_audio_s_per_s": world * 180.0 * both["steps_per_s_per_rank"],
                      "note": "config #2's e2e leg cannot exceed e2e_ceiling_audio_s_per_s on this host: it is "
                              "what the same bytes cost with no kernels at all"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
