import sys, torch
sys.path.insert(0, ".")
from aas_enhancement_b200 import LMFBFrontEnd
dev = torch.device("cuda", 0)
for n, samples in ((256, 160000), (30, 96000)):
    tmax = 1 + samples // 160
    fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev)
    wave = (0.1 * torch.randn(n, samples, device=dev)).clamp_(-1, 1)
    lens = torch.full((n,), samples, dtype=torch.int32, device=dev)
    mr = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
    mi = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
    g = torch.randn(n, 40, tmax, device=dev)
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
        print("n=%d %gs  backward%s: %.3f ms" % (n, samples / 16000, " + grad_wave" if want else "", e0.elapsed_time(e1) / 20))
