#!/usr/bin/env python3
"""BASELINE.json configs[4]: front-end throughput over batch size x utterance length, at 1/2/4/8 GPUs
(utterance-sharded weak scaling: the batch column is PER GPU), next to the host-CPU reference.

    python tools/sweep.py [out.md]                                    # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \\
           --master-port 29511 tools/sweep.py out.md                  # 8 GPUs, one rank each

GPU column: kernel-resident numbers (CUDA-graph replay of the C-ABI forward + backward, inputs in HBM,
ring of batches larger than L2 where it fits, CUDA events, max over ranks).  CPU column: the reference's
feature path restated with torch CPU ops (oracle/lmfb_torch_cpu.py), batched, all host threads, on a
bounded sample (<= 32 utterances of that length; rank 0 only), median of 5.
"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                  # noqa: E402
import torch.distributed as dist              # noqa: E402
import bench                                  # noqa: E402
from aas_enhancement_b200 import LMFBFrontEnd, _lib   # noqa: E402

import json                                   # noqa: E402
CPU_CACHE = os.path.join(ROOT, "gpurun_out", "cpu_sweep_cache.json")
if "--cpu-only" in sys.argv:
    # the CPU column, measured by ONE process with the machine to itself (under torchrun the other ranks'
    # polling threads would share the cores); cached for the multi-GPU runs on the same box
    secs_list = (1, 6, 30) if "--quick" in sys.argv else (1, 2, 4, 6, 10, 15, 30)
    out = {}
    for secs in secs_list:
        n_s = 32 if secs <= 10 else 8
        step, _ = bench._cpu_step_fns(n_s, int(secs * bench.SR), False)
        torch.set_num_threads(os.cpu_count() or 1)
        t, _ = bench._median_time(step, warm=2, reps=5, budget_s=20.0)
        out[str(secs)] = n_s * secs / t
        print("cpu", secs, "s:", out[str(secs)], "audio-s/s")
    os.makedirs(os.path.dirname(CPU_CACHE), exist_ok=True)
    json.dump({"cores": os.cpu_count(), "audio_s_per_s": out}, open(CPU_CACHE, "w"))
    sys.exit(0)

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev)
peak, _ = bench.peaks()
quick = "--quick" in sys.argv
SECS = (1, 6, 30) if quick else (1, 2, 4, 6, 10, 15, 30)
BATCH = (1, 30, 256) if quick else (1, 2, 4, 8, 16, 32, 64, 128, 256, 512)
cpu_cache = {}


class _NoClock:
    def region(self, on):
        pass


def cpu_rate(secs):
    """audio-s/s of the batched CPU path at this utterance length (32 utterances, all threads)."""
    if not cpu_cache and os.path.exists(CPU_CACHE):
        cpu_cache.update({int(k): v for k, v in json.load(open(CPU_CACHE))["audio_s_per_s"].items()})
    if secs not in cpu_cache and world > 1:
        return float("nan")                     # (never time the CPU path next to seven other ranks)
    if secs not in cpu_cache:
        n_s = 32 if secs <= 10 else 8
        step, _ = bench._cpu_step_fns(n_s, int(secs * bench.SR), False)
        torch.set_num_threads(os.cpu_count() or 1)
        t, _ = bench._median_time(step, warm=1, reps=5, budget_s=15.0)
        cpu_cache[secs] = n_s * secs / t
    return cpu_cache[secs]


rows = []
for secs in SECS:
    for n in BATCH:
        samples = int(secs * bench.SR)
        tmax = 1 + samples // bench.HOP
        slot_bytes = (n * samples + 4 * n * 161 * tmax + 3 * n * 40 * tmax) * 4
        if slot_bytes > 24e9:
            continue
        name = f"sweep_{n}x{secs}s"
        bench.WORKLOADS[name] = (n, float(secs), "masked")
        runner = bench.Runner(lib, _lib, fe, name, dev, 123 + rank)
        for i in range(3):
            runner.step(i)
        torch.cuda.synchronize()
        units = bench.time_device(runner, runner.ring, world, dev, _NoClock(), min_region_s=0.05)
        ms = statistics.median(units) / runner.ring
        rows.append((secs, n, ms, world * n * secs / (ms / 1e3),
                     runner.frames * runner.step_bytes / (ms / 1e3) / 1e9 / peak, runner.ring * runner.slot_bytes / 1e6))
        del runner
        torch.cuda.empty_cache()

if rank == 0:
    out = ["| seconds | batch per GPU | ms / fwd+bwd | audio-s/s (%d GPU) | %% of HBM roofline per GPU (5,624 B/frame) | CPU audio-s/s | GPU / CPU | resident ring MB |" % world,
           "|---|---|---|---|---|---|---|---|"]
    for secs, n, ms, rate, frac, mb in rows:
        c = cpu_rate(secs)
        out.append("| %d | %d | %.4f | %.3e | %.1f | %.3e | %.0f | %.0f |" % (secs, n, ms, rate, 100 * frac, c, rate / c, mb))
    text = "\n".join(out) + "\n"
    print(text)
    paths = [a for a in sys.argv[1:] if not a.startswith("-")]
    if paths:
        cores = os.cpu_count()
        with open(paths[0], "w") as f:
            f.write("# %d x B200, 'reim' mask, per-bin CMVN, CUDA-graph replay, CUDA events, max over ranks; CPU column: "
                    "batched torch CPU path on %d host threads (tools/sweep.py)\n" % (world, cores) + text)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
