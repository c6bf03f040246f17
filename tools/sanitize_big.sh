#!/bin/bash
# compute-sanitizer over a launch with more tiles than resident CTAs (cluster-launch-control scheduling)
mkdir -p gpurun_out
cat > /tmp/sanb.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from aas_enhancement_b200 import LMFBFrontEnd
torch.manual_seed(0)
n, samples = 52, 80000
tmax = 1 + samples // 160
fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").cuda()
wave = (0.1 * torch.randn(n, samples, device="cuda")).clamp_(-1, 1)
lens = torch.randint(samples // 2, samples + 1, (n,), dtype=torch.int32, device="cuda")
mr = torch.rand(n, 161, tmax, device="cuda", requires_grad=True)
mi = torch.rand(n, 161, tmax, device="cuda", requires_grad=True)
z, fl = fe(wave, lens, mr, mi); z.backward(torch.randn_like(z))
torch.cuda.synchronize()
print("tiles", n * ((tmax + 31) // 32), float(z.abs().sum()), float(mr.grad.abs().sum()))
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=lmfb_k1 python /tmp/sanb.py 2>&1 | grep -E "tiles|SUMMARY|Race reported|Invalid|Error" | sort | uniq -c | head -12 | tee gpurun_out/sanitize_big_$tool.txt
done
