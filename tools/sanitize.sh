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
# this round's paths: int16 PCM waves, two channels, a dense (generic) basis, the waveform gradient
import numpy as np
b = _synth.make_batch(3, 9000, seed=6, ragged=True)
lengths = torch.from_numpy(b["lengths"]).cuda()
g = torch.from_numpy(b["grad_out"]).cuda()
def masks(rows=161):
    rs = np.random.RandomState(1)
    return (torch.from_numpy(rs.rand(3, rows, b["tmax"]).astype(np.float32)).cuda().requires_grad_(True),
            torch.from_numpy(rs.rand(3, rows, b["tmax"]).astype(np.float32)).cuda().requires_grad_(True))
fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").cuda()
pcm = torch.from_numpy(np.round(b["wave"] * 32768).clip(-32768, 32767).astype(np.int16)).cuda()
mr, mi = masks(); z, _ = fe(pcm, lengths, mr, mi); z.backward(g); print("int16", float(z.abs().sum()))
w2 = torch.from_numpy(np.stack([b["wave"], b["wave"][::-1].copy()], axis=1)).cuda()
mr, mi = masks(322); z, _ = fe(w2, lengths, mr, mi); z.backward(g); print("2ch", float(z.abs().sum()))
fd = LMFBFrontEnd(mel_basis=np.random.RandomState(2).rand(40, 161) * 0.02, mask_mode="reim", cmvn_mode="per_bin").cuda()
mr, mi = masks(); z, _ = fd(torch.from_numpy(b["wave"]).cuda(), lengths, mr, mi); z.backward(g); print("dense", float(z.abs().sum()))
wg = torch.from_numpy(b["wave"]).cuda().requires_grad_(True)
mr, mi = masks(); z, _ = fe(wg, lengths, mr, mi); z.backward(g); print("grad_wave", float(wg.grad.abs().sum()))
# rows of 813 frames: the block-per-row CMVN kernels (two sums behind one pair of barriers)
bl = _synth.make_batch(2, 130000, seed=7, ragged=True)
mrl = torch.from_numpy(bl["mask_r"]).cuda().requires_grad_(True); mil = torch.from_numpy(bl["mask_i"]).cuda().requires_grad_(True)
z, _ = fe(torch.from_numpy(bl["wave"]).cuda(), torch.from_numpy(bl["lengths"]).cuda(), mrl, mil)
z.backward(torch.from_numpy(bl["grad_out"]).cuda()); print("long rows", float(z.abs().sum()), float(mrl.grad.abs().sum()))
torch.cuda.synchronize()
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --kernel-regex 'kns=lmfb_k1|cmvn_' python /tmp/san.py 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|Saved host" | tail -25 | tee gpurun_out/sanitize_$tool.txt
done
