#!/usr/bin/env python3
"""LMFB front-end benchmark (BASELINE.json metric: LMFB fwd+bwd audio-seconds / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one synthetic batch:

  masked workloads   forward + backward into both masks ('reim' mask mode, per-bin CMVN)
                     chime4_30x6s (default, BASELINE.json configs[1]), cfg0_8x4s, sweep_*
  paired workloads   configs[3] (FSEGAN / minimize_DCE): masked NOISY forward + backward plus an
                     unmasked CLEAN forward of the same utterances (two front-end passes)
  aas_step_30x6s     configs[2]: the whole AAS training step around the front-end (enhancer,
                     discriminator, acoustic model, CTC, NCCL gradient all-reduce), see bench_aas.py

Under torchrun each rank owns its own batch ring (utterance-sharded, weak scaling, no data-path
collective).  Prints ONE JSON line (rank 0):

  value          inputs resident in HBM, CUDA-graph replay of the C-ABI calls.  EXACTLY --steps steps
                 are timed as a unit with CUDA events; the unit is repeated until the timed region is
                 at least 0.5 s long and the MEDIAN unit (max over ranks per repeat) is reported
  e2e            the public autograd API with pinned HOST buffers, H2D/D2H copies inside the timed
                 region (median run)
  roofline       the dominant kernel, CUDA events recorded by the library around that kernel
  roofline_large the same step on sweep_256x10s (1.9 GB resident, far beyond L2), timed in the same
                 run: this is the size the HBM-roofline claim is made on (config #2 is 100 MB and
                 latency-bound)
  cpu_baseline   the reference's CPU path (oracle/lmfb_torch_cpu.py) on this host: all threads,
                 one thread, and the per-utterance loop a DataLoader worker would run

`--impl reference` times that CPU path alone (rank 0), same config dictionary.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

HOP = 160
SR = 16000
N_MELS = 40
B_FWD, B_BWD = 2088, 3536             # algorithmic bytes / frame, 'reim' 
B_FWD_CLEAN = 800                     # unmasked forward: wave 640 + out 160
B_K1_BWD = 640 + 1288 + 1288          # K1-backward's own compulsory bytes: wave + masks + grad masks
B_K1_FWD = 640 + 1288 + 160
MIN_REGION_S = 0.5                    # the timed region is at least this long whatever --steps says

WORKLOADS = {
    # name: (utterances, seconds, kind)
    "chime4_30x6s": (30, 6.0, "masked"),       # BASELINE.json configs[1]
    "cfg0_8x4s": (8, 4.0, "masked"),           # configs[0]
    "sweep_512x30s": (512, 30.0, "masked"),    # largest single-GPU point of configs[4]
    "sweep_256x10s": (256, 10.0, "masked"),
    "paired_30x6s": (30, 6.0, "paired"),       # configs[3]
    "paired_256x10s": (256, 10.0, "paired"),
    "aas_step_30x6s": (30, 6.0, "aas_step"),   # configs[2]
}
LARGE = "sweep_256x10s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ring_for(n, secs, kind):
    """Distinct resident batches cycled step by step, so that every step reads inputs that are not
    in L2 (>= ~400 MB > 126 MB)."""
    samples = int(secs * SR)
    tmax = 1 + samples // HOP
    slot_bytes = (n * samples + 4 * n * 161 * tmax + 3 * n * N_MELS * tmax) * 4
    if kind == "paired":
        slot_bytes += (n * samples + n * N_MELS * tmax) * 4
    ring = max(2, min(16, -(-400_000_000 // slot_bytes)))
    if slot_bytes * ring > 60e9:
        ring = max(1, int(60e9 // slot_bytes))
    return ring, slot_bytes


def workload_config(name, world):
    """The `config` object of the JSON line: identical for both arms (ours / reference)."""
    n, secs, kind = WORKLOADS[name]
    samples = int(secs * SR)
    tmax = 1 + samples // HOP
    ring, slot_bytes = ring_for(n, secs, kind)
    what = {"masked": "forward + backward into both masks",
            "paired": "noisy: masked forward + backward; clean: unmasked forward (two front-end passes, "
                      "the pair of utterances counted once)",
            "aas_step": "full AAS training step (G, D, ASR + CTC, front-end inside, gradient all-reduce)"}[kind]
    return {"workload": name, "utterances_per_gpu": n, "seconds": secs,
            "frames_per_step_per_gpu": n * tmax, "mask_mode": "reim", "cmvn": "per_bin", "step": what,
            "sharding": f"utterance-sharded x{world}, no data-path collective",
            "l2": f"inputs larger than L2: ring of {ring} distinct resident batches "
                  f"({ring * slot_bytes / 1e6:.0f} MB) cycled step by step"}


def pin_to_gpu_numa_node(index):
    """Run this rank (and allocate its pinned host buffers, first touch) on the CPUs next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls NVML for SM clock / throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._in_region = False
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((self._in_region, mhz))
                if self._in_region:
                    for k, bit in names.items():
                        if mask & bit:
                            self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()

    def region(self, on):
        self._in_region = on

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        inside = [m for r, m in self.samples if r]
        return {"sm_mhz": statistics.median(inside) if inside else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples_in_region": len(inside)}


# ------------------------------------------------------------------------------- ours
class Slot:
    """One resident synthetic batch + its output buffers (device)."""

    def __init__(self, n, samples, tmax, n_mels, gen, dev, paired=False, row_pad=0):
        self.wave = (0.1 * torch.randn(n, samples, generator=gen, device=dev)).clamp_(-1, 1)
        self.lengths = torch.full((n,), samples, dtype=torch.int32, device=dev)
        # row_pad (experiment only, --pad-rows): mask rows padded to a multiple of 32 frames, i.e.
        # 128-byte aligned row segments; the default is the reference's contiguous (N, 161, T)
        tp = -(-tmax // row_pad) * row_pad if row_pad else tmax
        self.mr = torch.rand(n, 161, tp, generator=gen, device=dev)[:, :, :tmax]
        self.mi = torch.rand(n, 161, tp, generator=gen, device=dev)[:, :, :tmax]
        self.gout = torch.randn(n, n_mels, tmax, generator=gen, device=dev)
        self.out = torch.empty(n, n_mels, tmax, device=dev)
        self.stats = torch.empty(n, n_mels, 2, device=dev)
        self.gr = torch.empty(n, 161, tp, device=dev)[:, :, :tmax]
        self.gi = torch.empty(n, 161, tp, device=dev)[:, :, :tmax]
        self.ws = torch.empty(n, n_mels, tmax, device=dev)
        if paired:                                  # the clean stream of the same utterances
            self.wave_c = (0.1 * torch.randn(n, samples, generator=gen, device=dev)).clamp_(-1, 1)
            self.out_c = torch.empty(n, n_mels, tmax, device=dev)
            self.stats_c = torch.empty(n, n_mels, 2, device=dev)


class Runner:
    """A workload resident on one GPU: ring of batches + the C-ABI calls of one step."""

    def __init__(self, lib, _lib, fe, name, dev, seed, row_pad=0):
        self.lib, self._lib, self.fe, self.name = lib, _lib, fe, name
        self.n, self.secs, self.kind = WORKLOADS[name]
        self.paired = self.kind == "paired"
        self.samples = int(self.secs * SR)
        self.tmax = 1 + self.samples // HOP
        self.frames = self.n * self.tmax
        self.audio_s = self.n * self.secs
        self.n_mels = fe.plan.n_mels
        self.flags = _lib.MASK_MODES["reim"] | _lib.CMVN_MODES["per_bin"]
        self.flags_clean = _lib.MASK_MODES["none"] | _lib.CMVN_MODES["per_bin"]
        self.ring, self.slot_bytes = ring_for(self.n, self.secs, self.kind)
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)                       # config.py:62 default seed (+ rank)
        self.slots = [Slot(self.n, self.samples, self.tmax, self.n_mels, gen, dev, self.paired, row_pad)
                      for _ in range(self.ring)]
        self.tables = None if os.environ.get("AAS_BENCH_NO_TABLES") else fe.plan.tables(dev)   # (env: A/B of the table upload)
        self.step_bytes = B_FWD + B_BWD + (B_FWD_CLEAN if self.paired else 0)
        self.launches_per_step = None               # counted from the first profiled step (see count_launches)

    def _io(self, s, flags, wave, masked, out, stats, prof, backward=False):
        import ctypes
        st = torch.cuda.current_stream().cuda_stream
        kw = dict(flags=flags, n=self.n, n_ch=1, tmax=self.tmax, eps=0.0, wave=wave.data_ptr(),
                  wave_stride=wave.stride(0), wave_len=wave.shape[1], lengths=s.lengths.data_ptr(),
                  window=self.fe.window.data_ptr(), mel_dev=self.fe.mel_basis.data_ptr(),
                  out=out.data_ptr(), stats=stats.data_ptr(), tables=self.tables, cuda_stream=st,
                  prof=ctypes.cast(prof, ctypes.c_void_p) if prof is not None else None)
        if masked:
            kw.update(mask_r=s.mr.data_ptr(), mask_i=s.mi.data_ptr(), mask_stride_n=s.mr.stride(0),
                      mask_stride_f=s.mr.stride(1))
        if backward:
            kw.update(grad_out=s.gout.data_ptr(), grad_mask_r=s.gr.data_ptr(), grad_mask_i=s.gi.data_ptr(),
                      workspace=s.ws.data_ptr())
        return self._lib.make_io(**kw)

    def fwd(self, s, prof=None):
        import ctypes
        io = self._io(s, self.flags, s.wave, True, s.out, s.stats, prof)
        self._lib.check(self.lib.aas_lmfb_forward_ex(self.fe.plan.handle, ctypes.byref(io)))

    def fwd_clean(self, s):
        import ctypes
        io = self._io(s, self.flags_clean, s.wave_c, False, s.out_c, s.stats_c, None)
        self._lib.check(self.lib.aas_lmfb_forward_ex(self.fe.plan.handle, ctypes.byref(io)))

    def bwd(self, s, prof=None):
        import ctypes
        io = self._io(s, self.flags, s.wave, True, s.out, s.stats, prof, backward=True)
        self._lib.check(self.lib.aas_lmfb_backward_ex(self.fe.plan.handle, ctypes.byref(io)))

    def step(self, i, prof_f=None, prof_b=None):
        s = self.slots[i % self.ring]
        self.fwd(s, prof_f)
        if self.paired:
            self.fwd_clean(s)
        self.bwd(s, prof_b)


def count_kernel_launches(fn):
    """Kernels one call of `fn` launches, counted with the CUDA profiler activity API (torch.profiler)."""
    try:
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as p:
            fn()
            torch.cuda.synchronize()
        names = [e.name for e in p.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        mine = [x for x in names if "aas_lmfb" in x or "lmfb_k1" in x or "cmvn" in x or "l1_abs" in x]
        return len(mine), sorted(set(mine))
    except Exception:
        return None, []


def time_device(runner, steps, world, dev, sampler, min_region_s=MIN_REGION_S):
    """EXACTLY `steps` steps timed as one unit (CUDA-graph replays, CUDA events); the unit is repeated
    until the region is >= min_region_s; returns per-repeat unit times in ms (max over ranks)."""
    import torch.distributed as dist
    ring = runner.ring
    side = torch.cuda.Stream()

    def capture(count):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(count):
                runner.step(i)
        return g

    reps, rem = divmod(steps, ring)
    g_full = capture(ring) if reps else None
    g_rem = capture(rem) if rem else None

    def unit():
        for _ in range(reps):
            g_full.replay()
        if rem:
            g_rem.replay()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    unit()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    unit()
    e1.record()
    torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1), 1e-3)
    repeats = int(min(4000, max(5, -(-min_region_s * 1e3 // est))))
    if world > 1:                                                    # every rank must time the same count
        t = torch.tensor([repeats], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        repeats = int(t.item())
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(repeats + 1)]
    sync_all()
    sampler.region(True)
    evs[0].record()
    for r in range(repeats):
        unit()
        evs[r + 1].record()
    torch.cuda.synchronize()
    sampler.region(False)
    times = torch.tensor([evs[r].elapsed_time(evs[r + 1]) for r in range(repeats)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.barrier()
    return times.cpu().tolist()


def time_kernels(runner, n_prof):
    """Per-kernel durations: CUDA events recorded by the library around each kernel, on the launch
    stream.  All profiled steps are enqueued back to back (the GPU never idles between them, as in
    the timed region); with a host sync after every step the event pairs of these 5-30 us kernels
    would mostly measure the CPU's launch latency."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(8)] for _ in range(n_prof)]
    for row in ev:
        for e in row:
            e.record()                                              # materialise the handles
    torch.cuda.synchronize()
    profs = [(ctypes_array([e.cuda_event for e in row[:4]]), ctypes_array([e.cuda_event for e in row[4:]])) for row in ev]
    for i in range(3):                                              # fill the launch queue ahead of the profiled steps
        runner.step(i)
    for i in range(n_prof):
        runner.step(i, profs[i][0], profs[i][1])
    torch.cuda.synchronize()
    k_ms = {"k1_fwd": [], "k2_fwd": [], "k2_bwd": [], "k1_bwd": []}
    for row in ev[n_prof // 4:]:                                    # the first quarter still overlaps the queue fill
        k_ms["k1_fwd"].append(row[0].elapsed_time(row[1]))
        k_ms["k2_fwd"].append(row[2].elapsed_time(row[3]))
        k_ms["k1_bwd"].append(row[4].elapsed_time(row[5]))
        k_ms["k2_bwd"].append(row[6].elapsed_time(row[7]))
    return {k: statistics.median(v) for k, v in k_ms.items()}


def roofline_block(runner, ms_per_step, k_ms, peak, peak_src):
    frames = runner.frames
    dom = "k1_bwd" if k_ms["k1_bwd"] >= k_ms["k1_fwd"] else "k1_fwd"
    dom_bytes = B_K1_BWD if dom == "k1_bwd" else B_K1_FWD
    achieved = frames * dom_bytes / (k_ms[dom] / 1e3) / 1e9
    step_gbs = frames * runner.step_bytes / (ms_per_step / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(runner.name, {}).get(dom + "_dram_bytes")
        except Exception:
            traffic = None
    return {"bound": "hbm", "workload": runner.name,
            "kernel": "lmfb_k1<reim,%s>" % ("bwd" if dom == "k1_bwd" else "fwd"),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": frames * dom_bytes,
            "avg_launch_ms": k_ms[dom],
            "kernels_ms": k_ms,
            "k1_fwd_frac": frames * B_K1_FWD / (k_ms["k1_fwd"] / 1e3) / 1e9 / peak,
            "k1_bwd_frac": frames * B_K1_BWD / (k_ms["k1_bwd"] / 1e3) / 1e9 / peak,
            "ms_per_step": ms_per_step,
            "step_algorithmic_bytes_per_frame": runner.step_bytes,
            "step_algorithmic_gbs_per_gpu": step_gbs, "step_frac": step_gbs / peak}


def run_ours(args):
    import torch.distributed as dist
    from aas_enhancement_b200 import LMFBFrontEnd, _lib, build as _build

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the LMFB path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = pin_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.lib:                                         # development A/B: another build of the library
        _lib.set_library_path(os.path.abspath(args.lib))
    elif rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    lib = _lib.load()

    if WORKLOADS[args.workload][2] == "aas_step":
        import bench_aas
        return bench_aas.run(args, world, rank, local, dev)

    fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev)
    fe.set_tuning(args.warps_fwd, args.warps_bwd, args.sched == "static")
    runner = Runner(lib, _lib, fe, args.workload, dev, 123 + rank, args.pad_rows)

    # ---- warm-up (also sets function attributes before any capture)
    warm = max(args.warmup, 3)
    for i in range(warm):
        runner.step(i)
    torch.cuda.synchronize()
    launches, launch_names = count_kernel_launches(lambda: runner.step(0))

    sampler = ClockSampler(local)
    sampler.start()
    units = time_device(runner, args.steps, world, dev, sampler)
    k_ms = time_kernels(runner, min(max(args.steps, 10), 100))

    # ---- the same step at the size the roofline claim is made on (1.9 GB resident)
    large = None
    if not args.no_large and args.workload != LARGE and not args.pad_rows:
        big = Runner(lib, _lib, fe, LARGE, dev, 1123 + rank)
        for i in range(3):
            big.step(i)
        torch.cuda.synchronize()
        big_units = time_device(big, 4 * big.ring, world, dev, sampler, min_region_s=0.3)
        big_k = time_kernels(big, 40)
        large = (big, statistics.median(big_units) / (4 * big.ring), big_k, len(big_units))
        del big.slots
        torch.cuda.empty_cache()
    clocks = sampler.stop()

    # ---- e2e: public autograd API, pinned host buffers, copies inside the timed region
    if args.no_e2e:                                          # development A/B runs only: not a bench line
        e2e = {"value": float("nan"), "unit": "audio-s/s", "skipped": True}
    else:
        e2e = measure_e2e(fe, runner, dev, min(max(args.steps, 20), 200), world)
        e2e["numa_local_cpus"] = numa_cpus

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    unit_ms = statistics.median(units)
    ms_per_step = unit_ms / args.steps
    value = world * runner.audio_s / (ms_per_step / 1e3)
    line = {
        "metric": "LMFB fwd+bwd audio-seconds per second",
        "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "timing": {"launch": "CUDA graph replay of the C-ABI calls; CUDA events; max over ranks per repeat",
                   "unit": "exactly --steps steps", "repeats": len(units),
                   "region_ms": sum(units), "unit_ms_median": unit_ms, "unit_ms_min": min(units),
                   "unit_ms_p90": sorted(units)[int(0.9 * (len(units) - 1))]},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": (launches if launches is not None else 3) * args.steps,
        "gpu_launches_per_step": launches, "gpu_kernels": launch_names,
        "roofline": roofline_block(runner, ms_per_step, k_ms, peak, peak_src),
    }
    if large is not None:
        big, big_ms, big_k, big_reps = large
        line["roofline_large"] = roofline_block(big, big_ms, big_k, peak, peak_src)
        line["roofline_large"]["value_audio_s_per_s"] = big.audio_s / (big_ms / 1e3)
        line["roofline_large"]["repeats"] = big_reps
    if args.no_e2e or args.pad_rows or args.lib or args.warps_fwd or args.warps_bwd or args.sched != "clc":
        line["experiment"] = ("not a bench line: " + ", ".join(
            x for x in ("end-to-end leg skipped" if args.no_e2e else "",
                        "library %s" % args.lib if args.lib else "",
                        "tuning knobs fwd=%d bwd=%d sched=%s" % (args.warps_fwd, args.warps_bwd, args.sched)
                        if (args.warps_fwd or args.warps_bwd or args.sched != "clc") else "",
                        "mask rows padded to %d frames (not the reference layout)" % args.pad_rows if args.pad_rows else "") if x))
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(runner.n, runner.samples, runner.paired)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ctypes_array(handles):
    import ctypes
    arr = (ctypes.c_void_p * len(handles))(*[ctypes.c_void_p(int(h)) for h in handles])
    return arr


def measure_e2e(fe, runner, dev, steps, world):
    """End to end through the public autograd API with HOST buffers: every step copies wave,
    lengths, both masks and grad_out (paired: also the clean wave) from pinned host memory, runs the
    step, and copies the features and both mask gradients (paired: also the clean features) back to
    pinned host memory.  Copies and compute of consecutive steps overlap on three streams
    (double-buffered), as a data loader would."""
    import torch.distributed as dist
    from aas_enhancement_b200 import LMFBFrontEnd
    n, samples, tmax, n_mels, paired = runner.n, runner.samples, runner.tmax, runner.n_mels, runner.paired
    fe_clean = LMFBFrontEnd(mask_mode="none", cmvn_mode="per_bin").to(dev) if paired else None
    gen = torch.Generator()
    gen.manual_seed(123)
    nbuf = int(os.environ.get("AAS_BENCH_E2E_NBUF", "2"))      # buffers in flight (experiment knob)
    host_in = []
    for _ in range(nbuf):
        d = dict(
            wave=(0.1 * torch.randn(n, samples, generator=gen)).clamp_(-1, 1).pin_memory(),
            lens=torch.full((n,), samples, dtype=torch.int32).pin_memory(),
            mr=torch.rand(n, 161, tmax, generator=gen).pin_memory(),
            mi=torch.rand(n, 161, tmax, generator=gen).pin_memory(),
            g=torch.randn(n, n_mels, tmax, generator=gen).pin_memory())
        if paired:
            d["wave_c"] = (0.1 * torch.randn(n, samples, generator=gen)).clamp_(-1, 1).pin_memory()
        host_in.append(d)
    host_out = []
    for _ in range(nbuf):
        d = dict(z=torch.empty(n, n_mels, tmax).pin_memory(),
                 gr=torch.empty(n, 161, tmax).pin_memory(),
                 gi=torch.empty(n, 161, tmax).pin_memory())
        if paired:
            d["zc"] = torch.empty(n, n_mels, tmax).pin_memory()
        host_out.append(d)
    dev_in = [{k: torch.empty_like(v, device=dev) for k, v in host_in[0].items()} for _ in range(nbuf)]
    h2d = sum(t.numel() * t.element_size() for t in host_in[0].values())
    d2h = sum(t.numel() * t.element_size() for t in host_out[0].values())
    s_in, s_c, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in = [torch.cuda.Event() for _ in range(nbuf)]
    ev_c = [torch.cuda.Event() for _ in range(nbuf)]
    ev_free = [torch.cuda.Event() for _ in range(nbuf)]
    ev_done = [torch.cuda.Event() for _ in range(nbuf)]
    keep = [None] * nbuf

    def one(i):
        b = i % nbuf
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_free[b])                      # compute of step i-nbuf has consumed the buffer
            for k in dev_in[b]:
                dev_in[b][k].copy_(host_in[b][k], non_blocking=True)
            ev_in[b].record(s_in)
        with torch.cuda.stream(s_c):
            s_c.wait_event(ev_in[b])
            s_c.wait_event(ev_done[b])                       # D2H of step i-nbuf has read the old results
            mr = dev_in[b]["mr"].detach().requires_grad_(True)
            mi = dev_in[b]["mi"].detach().requires_grad_(True)
            z, _ = fe(dev_in[b]["wave"], dev_in[b]["lens"], mr, mi)
            zc = None
            if paired:
                with torch.no_grad():
                    zc, _ = fe_clean(dev_in[b]["wave_c"], dev_in[b]["lens"])
            z.backward(dev_in[b]["g"])
            keep[b] = (z, mr, mi, zc)
            ev_c[b].record(s_c)
            ev_free[b].record(s_c)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_c[b])
            z, mr, mi, zc = keep[b]
            for t in (z, mr.grad, mi.grad) + ((zc,) if zc is not None else ()):
                t.record_stream(s_out)
            host_out[b]["z"].copy_(z.detach(), non_blocking=True)
            host_out[b]["gr"].copy_(mr.grad, non_blocking=True)
            host_out[b]["gi"].copy_(mi.grad, non_blocking=True)
            if zc is not None:
                host_out[b]["zc"].copy_(zc, non_blocking=True)
            ev_done[b].record(s_out)

    for i in range(24):                                      # allocator / autograd paths settle slowly
        one(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    runs = []
    t_begin = time.perf_counter()
    # 5..15 timed runs of `steps` steps (stop after ~4 s); the MEDIAN run is reported.  On a freshly
    # started box the first runs are sometimes several times slower on the HOST side (the image is
    # still paging in), which the median absorbs.
    n_runs = 0
    while n_runs < 5 or (n_runs < 15 and time.perf_counter() - t_begin < 4.0):
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_in)
        for i in range(steps):
            one(i)
        s_out.wait_stream(s_c)
        e1.record(s_out)
        torch.cuda.synchronize()
        runs.append((max(e0.elapsed_time(e1), 0.0), (time.perf_counter() - t0) * 1e3))
        n_runs += 1
        if world > 1:                                         # keep the ranks' run counts identical
            flag = torch.tensor([1 if (n_runs < 5 or (n_runs < 15 and time.perf_counter() - t_begin < 4.0)) else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0 and n_runs >= 5:
                break
    runs.sort()
    ms, wall_ms = runs[len(runs) // 2]
    ms_best = runs[0][0]
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * runner.audio_s * steps / (ms / 1e3), "unit": "audio-s/s",
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
            "runs": len(runs), "ms_median_run": ms, "ms_best_run": ms_best, "wall_ms": wall_ms,
            "h2d_gbs_per_rank": h2d * steps / (ms / 1e3) / 1e9, "d2h_gbs_per_rank": d2h * steps / (ms / 1e3) / 1e9,
            "api": "LMFBFrontEnd.forward + autograd backward; wave, lengths, both masks and grad_out "
                   "copied from pinned host memory, features and both mask gradients copied back to "
                   "pinned host memory, every step; copies of neighbouring steps overlap on 3 streams; "
                   "median of the timed runs, max over ranks"}


# ------------------------------------------------------------------------------- CPU arm
def _cpu_inputs(n, samples, seed=123):
    from oracle import lmfb_oracle as orc
    gen = torch.Generator()
    gen.manual_seed(seed)
    tmax = 1 + samples // HOP
    wave = (0.1 * torch.randn(n, samples, generator=gen)).clamp_(-1, 1)
    wave_c = (0.1 * torch.randn(n, samples, generator=gen)).clamp_(-1, 1)
    mr = torch.rand(n, 161, tmax, generator=gen)
    mi = torch.rand(n, 161, tmax, generator=gen)
    g = torch.randn(n, N_MELS, tmax, generator=gen)
    mel = torch.from_numpy(orc.mel_filterbank().astype(np.float32))
    win = torch.from_numpy(orc.hamming_window().astype(np.float32))
    return wave, wave_c, mr, mi, g, mel, win


def _cpu_step_fns(n, samples, paired):
    """(batched step, per-utterance-loop step) of the reference's CPU path on seeded inputs."""
    from oracle import lmfb_torch_cpu as cpu
    wave, wave_c, mr, mi, g, mel, win = _cpu_inputs(n, samples)
    lengths = [samples] * n

    def batched():
        cpu.fwd_bwd_batched(wave, mr, mi, g, mel, win, "per_bin", "reim")
        if paired:
            with torch.no_grad():
                cpu.forward_batched(wave_c, None, None, mel, win, "per_bin", "none")

    def per_utt():
        cpu.fwd_bwd_per_utterance(wave, lengths, mr, mi, g, mel, win, "per_bin")
        if paired:
            with torch.no_grad():
                for i in range(n):
                    cpu.forward_batched(wave_c[i:i + 1], None, None, mel, win, "per_bin", "none")
    return batched, per_utt


def _median_time(fn, warm=3, reps=10, budget_s=20.0):
    for _ in range(warm):
        fn()
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < reps and (len(times) < 3 or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return statistics.median(times), len(times)


def cpu_sample_size(n):
    return min(n, 32)                                  # bounded sample of the workload (utterances)


def cpu_baseline(n, samples, paired=False):
    """The oracle-side torch CPU path (kind 'port'), BASELINE.md section 2 protocol: seeded inputs,
    3 warm-ups, median of 10; batched with all threads and with one thread, and the per-utterance
    loop (how the reference's DataLoader worker would run it)."""
    cores = os.cpu_count() or 1
    n_s = cpu_sample_size(n)
    batched, per_utt = _cpu_step_fns(n_s, samples, paired)
    audio = n_s * samples / SR
    torch.set_num_threads(cores)
    t_all, r_all = _median_time(batched)
    t_loop, r_loop = _median_time(per_utt, budget_s=10.0)
    torch.set_num_threads(1)
    t_one, r_one = _median_time(batched, warm=1, reps=5, budget_s=10.0)
    torch.set_num_threads(cores)
    model = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"value": audio / t_all, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "cpu_model": model,
            "threads_1_value": audio / t_one,
            "per_utterance_loop_value": audio / t_loop,
            "sample": f"{n_s} x {samples / SR:g} s per pass (oracle/lmfb_torch_cpu.py: torch.stft + "
                      f"model.py:191-198 ops + CMVN + autograd{'; + unmasked clean forward' if paired else ''}); "
                      f"median of {r_all} batched passes on {cores} threads {t_all * 1e3:.1f} ms, "
                      f"of {r_one} on 1 thread {t_one * 1e3:.1f} ms, "
                      f"of {r_loop} per-utterance loops on {cores} threads {t_loop * 1e3:.1f} ms"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n, secs, kind = WORKLOADS[args.workload]
    if kind == "aas_step":
        import bench_aas
        return bench_aas.run_reference(args, world)
    samples = int(secs * SR)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = cpu_sample_size(n)                            # fixed sample: all 30 utterances of config #2
    step, _ = _cpu_step_fns(n_s, samples, kind == "paired")
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    value = n_s * secs / med
    sample = (f"{n_s} of {n} utterances x {secs:g} s per step, batched torch CPU path "
              f"(oracle/lmfb_torch_cpu.py), {torch.get_num_threads()} threads, median of {len(times)} steps "
              f"(min {min(times) * 1e3:.1f} ms, max {max(times) * 1e3:.1f} ms)")
    line = {
        "impl": "reference",
        "metric": "LMFB fwd+bwd audio-seconds per second", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="chime4_30x6s", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-large", action="store_true", help="skip the roofline_large leg (sweep_256x10s in the same run)")
    ap.add_argument("--pad-rows", type=int, default=0, help="experiment: pad the mask rows to a multiple of this many frames (not the reference layout; not a valid bench line)")
    ap.add_argument("--lib", default="", help="experiment: load this build of libaas_lmfb.so")
    ap.add_argument("--warps-fwd", type=int, default=0, help="experiment: warps per tile of the forward kernel (2..5)")
    ap.add_argument("--warps-bwd", type=int, default=0, help="experiment: warps per tile of the backward kernel (2..5)")
    ap.add_argument("--sched", default="clc", choices=["clc", "static"], help="experiment: tile schedule")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (kernel A/B runs; not a valid bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 20
        args.warmup = args.warmup if args.warmup is not None else 3
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else 2000
        args.warmup = args.warmup if args.warmup is not None else 20
        run_ours(args)


if __name__ == "__main__":
    main()
