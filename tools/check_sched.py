#!/usr/bin/env python3
"""GPU: the dynamic (cluster-launch-control) and static tile schedules give bit-identical results on
a launch with more tiles than resident CTAs.  usage: python tools/check_sched.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from aas_enhancement_b200 import LMFBFrontEnd
torch.manual_seed(0)
dev = torch.device("cuda", 0)
n, samples = 96, 80000
tmax = 1 + samples // 160
fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev).set_tuning(0, 0, sys.argv[2] == "static")
wave = (0.1 * torch.randn(n, samples, device=dev)).clamp_(-1, 1)
lens = torch.randint(samples // 2, samples + 1, (n,), dtype=torch.int32, device=dev)
mr = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
mi = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
g = torch.randn(n, 40, tmax, device=dev)
z, fl = fe(wave, lens, mr, mi); z.backward(g)
torch.cuda.synchronize()
torch.save({"z": z.detach().cpu(), "gr": mr.grad.cpu(), "gi": mi.grad.cpu()}, sys.argv[1])
print("ok", float(z.abs().sum()))
''' % ROOT
outs = []
for sched in ("static", "clc"):
    path = "/tmp/sched_%s.pt" % sched
    subprocess.check_call(["timeout", "120", sys.executable, "-c", CODE, path, sched])
    outs.append(path)
import torch
a, b = torch.load(outs[0]), torch.load(outs[1])
for k in a:
    same = torch.equal(a[k], b[k])
    print(k, "identical" if same else "DIFFERENT max|d| %g" % float((a[k] - b[k]).abs().max()))
    assert same and torch.isfinite(a[k]).all()
print("schedules agree")
