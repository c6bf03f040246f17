#!/usr/bin/env python3
"""Forward time per mask mode (and the STFT op) on 256 x 10 s and 30 x 6 s: the clean stream of the
AAS / FSEGAN trainers runs the unmasked forward (BASELINE.json configs[3])."""
import sys, torch
sys.path.insert(0, ".")
from aas_enhancement_b200 import LMFBFrontEnd
dev = torch.device("cuda", 0)
for n, samples in ((256, 160000), (30, 96000)):
    tmax = 1 + samples // 160
    wave = (0.1 * torch.randn(n, samples, device=dev)).clamp_(-1, 1)
    lens = torch.full((n,), samples, dtype=torch.int32, device=dev)
    mr = torch.rand(n, 161, tmax, device=dev); mi = torch.rand(n, 161, tmax, device=dev)
    for mode in ("reim", "power", "none", "stft"):
        fe = LMFBFrontEnd(mask_mode=mode if mode != "stft" else "none", cmvn_mode="per_bin").to(dev)
        def run():
            with torch.no_grad():
                if mode == "stft": return fe.stft(wave, lens)
                return fe(wave, lens, mr if mode != "none" else None, mi if mode == "reim" else None)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("n=%d %gs  forward %-5s: %.3f ms  %.3e audio-s/s" % (n, samples / 16000, mode, ms, n * samples / 16000 / (ms / 1e3)))
