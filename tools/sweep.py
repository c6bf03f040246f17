#!/usr/bin/env python3
"""BASELINE.json configs[4]: front-end throughput over batch size x utterance length on ONE GPU
(kernel-resident numbers: CUDA-graph replay of the C-ABI forward+backward, inputs in HBM, ring of
batches larger than L2 where it fits).  Writes a markdown table.   usage: tools/sweep.py [out.md]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aas_enhancement_b200 import LMFBFrontEnd, _lib

dev = torch.device("cuda", 0)
lib = _lib.load()
fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev)
flags = _lib.MASK_MODES["reim"] | _lib.CMVN_MODES["per_bin"]
PEAK = 6452.8e9
B_STEP = 5624.0
rows = []
for secs in (1, 2, 4, 6, 10, 15, 30):
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
        samples = secs * 16000
        tmax = 1 + samples // 160
        slot_bytes = (n * samples + 4 * n * 161 * tmax + 3 * n * 40 * tmax) * 4
        if slot_bytes > 24e9:
            continue
        ring = max(1, min(8, int(400e6 // slot_bytes) + 1))
        if slot_bytes * ring > 40e9:
            ring = 1
        slots = []
        for _ in range(ring):
            s = dict(wave=(0.1 * torch.randn(n, samples, device=dev)).clamp_(-1, 1),
                     lengths=torch.full((n,), samples, dtype=torch.int32, device=dev),
                     mr=torch.rand(n, 161, tmax, device=dev), mi=torch.rand(n, 161, tmax, device=dev),
                     g=torch.randn(n, 40, tmax, device=dev), out=torch.empty(n, 40, tmax, device=dev),
                     stats=torch.empty(n, 40, 2, device=dev), ws=torch.empty(n, 40, tmax, device=dev))
            s["gr"], s["gi"] = torch.empty_like(s["mr"]), torch.empty_like(s["mi"])
            slots.append(s)

        def step(s):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.aas_lmfb_forward(fe.plan.handle, s["wave"].data_ptr(), s["lengths"].data_ptr(), n, s["wave"].stride(0),
                                            s["mr"].data_ptr(), s["mi"].data_ptr(), s["mr"].stride(0), s["mr"].stride(1),
                                            fe.window.data_ptr(), s["out"].data_ptr(), s["stats"].data_ptr(), tmax, flags, 0.0, st, None))
            _lib.check(lib.aas_lmfb_backward(fe.plan.handle, s["wave"].data_ptr(), s["lengths"].data_ptr(), n, s["wave"].stride(0),
                                             s["mr"].data_ptr(), s["mi"].data_ptr(), s["mr"].stride(0), s["mr"].stride(1),
                                             fe.window.data_ptr(), s["out"].data_ptr(), s["stats"].data_ptr(), s["g"].data_ptr(),
                                             s["gr"].data_ptr(), s["gi"].data_ptr(), s["ws"].data_ptr(), tmax, flags, 0.0, st, None))
        for s in slots:
            step(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.graph(g, stream=side):
            for s in slots:
                step(s)
        g.replay(); torch.cuda.synchronize()
        frames = n * tmax
        reps = max(3, min(200, int(0.05 / max(1e-6, ring * frames * B_STEP / 3e12))))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (reps * ring)
        rows.append((secs, n, ms, n * secs / (ms / 1e3), frames * B_STEP / (ms / 1e3) / PEAK, ring * slot_bytes / 1e6))
        del slots, g
        torch.cuda.empty_cache()
out = ["| seconds | batch | ms / fwd+bwd | audio-s/s | % of HBM roofline (5,624 B/frame) | resident ring MB |", "|---|---|---|---|---|---|"]
for r in rows:
    out.append("| %d | %d | %.4f | %.3e | %.1f | %.0f |" % (r[0], r[1], r[2], r[3], 100 * r[4], r[5]))
text = "\n".join(out) + "\n"
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write("# one B200, 'reim' mask, per-bin CMVN, CUDA-graph replay, CUDA events (tools/sweep.py)\n" + text)
