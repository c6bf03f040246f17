import sys, time, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench
from aas_enhancement_b200 import LMFBFrontEnd
dev = torch.device("cuda", 0)
fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev)
n, samples = 30, 96000
tmax = 601
for trial in range(2):
    r = bench.measure_e2e(fe, n, samples, tmax, 40, 180.0, dev, 40, 1)
    print("pipelined", "%.3e" % r["value"], "wall_ms/step %.3f" % (r["wall_ms"] / 40))
# time pieces
import torch
h = torch.randn(n, 161, tmax).pin_memory(); d = torch.empty_like(h, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("h2d 11.6MB ms", (time.perf_counter() - t0) * 10)
wave = torch.randn(n, samples, device=dev); lens = torch.full((n,), samples, dtype=torch.int32, device=dev)
mr = torch.rand(n, 161, tmax, device=dev); mi = torch.rand(n, 161, tmax, device=dev); g = torch.randn(n, 40, tmax, device=dev)
for _ in range(3):
    a = mr.detach().requires_grad_(True); b = mi.detach().requires_grad_(True)
    z, _ = fe(wave, lens, a, b); z.backward(g)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(100):
    a = mr.detach().requires_grad_(True); b = mi.detach().requires_grad_(True)
    z, _ = fe(wave, lens, a, b); z.backward(g)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("fwd+bwd api: cpu ms/step %.3f total ms/step %.3f" % ((t1 - t0) * 10, (t2 - t0) * 10))
