#!/bin/bash
# compute-sanitizer over a small forward+backward (all mask modes, ragged lengths, boundary tiles).
# usage: tools/sanitize.sh   (on the GPU box)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import _synth
from aas_enhancement_b200 import LMFBFrontEnd
for mode in ("reim", "power", "none"):
    b = _synth.make_batch(3, 9000, seed=5, ragged=True)
    fe = LMFBFrontEnd(mask_mode=mode, cmvn_mode="per_bin").cuda()
    wave = torch.from_numpy(b["wave"]).cuda(); lengths = torch.from_numpy(b["lengths"]).cuda()
    mr = torch.from_numpy(b["mask_r"]).cuda().requires_grad_(mode != "none")
    mi = torch.from_numpy(b["mask_i"]).cuda().requires_grad_(mode == "reim")
    z, fl = fe(wave, lengths, mr if mode != "none" else None, mi if mode == "reim" else None)
    if mode != "none":
        z.backward(torch.from_numpy(b["grad_out"]).cuda())
    s, _ = fe.stft(wave, lengths)
    torch.cuda.synchronize()
    print(mode, float(z.abs().sum()), float(s.abs().sum()))
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=lmfb_k1 python /tmp/san.py 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|Saved host" | tail -25 | tee gpurun_out/sanitize_$tool.txt
done
