#!/usr/bin/env python3
"""Cost of the waveform gradient: backward through autograd with and without `wave.requires_grad`,
per warps-per-tile shape of the backward kernel (0 = the library's default)."""
import sys, torch
sys.path.insert(0, ".")
from aas_enhancement_b200 import LMFBFrontEnd
dev = torch.device("cuda", 0)
for n, samples in ((256, 160000), (30, 96000)):
    tmax = 1 + samples // 160
    wave = (0.1 * torch.randn(n, samples, device=dev)).clamp_(-1, 1)
    lens = torch.full((n,), samples, dtype=torch.int32, device=dev)
    mr = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
    mi = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
    g = torch.randn(n, 40, tmax, device=dev)
    for wb in (0, 4, 5, 8):
        fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev).set_tuning(0, wb)
        for want in (False, True):
            w = wave.clone().requires_grad_(want)
            z, _ = fe(w, lens, mr, mi)
            for _ in range(3):
                z.backward(g, retain_graph=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                z.backward(g, retain_graph=True)
            e1.record(); torch.cuda.synchronize()
            print("n=%d %gs warps_bwd=%d backward%s: %.3f ms" % (n, samples / 16000, wb, " + grad_wave" if want else "", e0.elapsed_time(e1) / 20))
