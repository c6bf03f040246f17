#!/usr/bin/env python3
"""Per-phase clock timeline of the K1 kernels (debug build with -DLMFB_TIMELINE into a scratch .so).
Run on the GPU box:  python tools/timeline.py [workload]"""
import os, subprocess, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aas_enhancement_b200 import build as B
# built in-tree (git-ignored *.so) so that a copy made on the CPU container travels to the GPU box:
#   python tools/timeline.py --build-only [-D...]
lib_dbg = os.path.join(ROOT, "aas_enhancement_b200", "libaas_lmfb_timeline%s.so" % os.environ.get("TL_TAG", ""))
extra = [a for a in sys.argv[1:] if a.startswith("-D")]
if "--build-only" in sys.argv or not os.path.exists(lib_dbg):
    subprocess.check_call([B.find_nvcc()] + B.NVCC_FLAGS + ["-DLMFB_TIMELINE"] + extra + [B.SRC, "-o", lib_dbg])
    print("debug build flags:", extra)
    if "--build-only" in sys.argv:
        sys.exit(0)
import torch
from aas_enhancement_b200 import _lib
_lib.LIB_PATH = lib_dbg
from aas_enhancement_b200 import LMFBFrontEnd
wl = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "sweep"
n, samples = (256, 160000) if wl == "sweep" else (30, 96000)
tmax = 1 + samples // 160
dev = torch.device("cuda", 0)
W = int(os.environ.get("TL_W", "5"))
buf_f = torch.zeros(8 * 64 * 8 * 8 + 4096, dtype=torch.int64, device=dev)
buf_b = torch.zeros(8 * 64 * 8 * 8 + 4096, dtype=torch.int64, device=dev)
os.environ["AAS_LMFB_TIMELINE_FWD"] = str(buf_f.data_ptr())
os.environ["AAS_LMFB_TIMELINE_BWD"] = str(buf_b.data_ptr())
fe = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev).set_tuning(W, W)
wave = (0.1 * torch.randn(n, samples, device=dev)).clamp_(-1, 1)
lens = torch.full((n,), samples, dtype=torch.int32, device=dev)
mr = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
mi = torch.rand(n, 161, tmax, device=dev, requires_grad=True)
g = torch.randn(n, 40, tmax, device=dev)
for _ in range(3):
    z, _ = fe(wave, lens, mr, mi); z.backward(g)
torch.cuda.synchronize()
names = ["stage", "wait1", "pass1+ld", "wait2", "pass2*", "p3A(neg)", "p3B*"]  # fwd: tl5 = phase-3 mid stamp
for label, buf in (("fwd", buf_f), ("bwd", buf_b)):
    first = buf.cpu()[8 * 64 * 8 * 8:].view(1024, 4)
    fv = first[first[:, 0] > 0]
    if fv.numel():
        end = fv[:, 0].double(); loop = fv[:, 1].double(); entry = fv[:, 2].double(); sm = fv[:, 3]
        k0 = entry.min()
        print(label, 'CTAs: n=%d on %d SMs | entry (us after first) mean %.1f max %.1f | prologue us mean %.2f max %.2f | end us mean %.1f min %.1f max %.1f'
              % (len(fv), len(set(sm.tolist())), (entry - k0).mean() / 1e3, (entry - k0).max() / 1e3, (loop - entry).mean() / 1e3, (loop - entry).max() / 1e3,
                 (end - k0).mean() / 1e3, (end - k0).min() / 1e3, (end - k0).max() / 1e3))
        import collections
        per_sm = collections.Counter(sm.tolist())
        print('   CTAs per SM: min %d max %d' % (min(per_sm.values()), max(per_sm.values())))
    t = buf.cpu()[:8 * 64 * 8 * 8].view(8, 64, -1)[:, :, :W * 8].reshape(8, 64, W, 8).double()
    valid = t[..., 7] > 0
    d = t[..., 1:] - t[..., :-1]
    print(label, "cycles per phase, mean over", int(valid.sum()), "warp-tiles; per warp index")
    for w in range(W):
        m = valid[:, :, w]
        row = [float(d[:, :, w, i][m].mean()) for i in range(7)]
        tot = float((t[:, :, w, 7] - t[:, :, w, 0])[m].mean())
        if label == "fwd":
            a = float((t[:, :, w, 5] - t[:, :, w, 6])[m].mean()); b = float((t[:, :, w, 7] - t[:, :, w, 5])[m].mean())
            p2 = float((t[:, :, w, 6] - t[:, :, w, 4])[m].mean())
            print(f"     fwd detail: pass2+barrier={p2:8.0f}  phase3 stageA={a:8.0f}  stageB={b:8.0f}")
        print("  warp", w, " ".join(f"{nm}={v:8.0f}" for nm, v in zip(names, row)), f" total={tot:9.0f}")
    # tile-to-tile period of one CTA
    per = (t[1:, :, 0, 0] - t[:-1, :, 0, 0])[valid[1:, :, 0] & valid[:-1, :, 0]]
    if per.numel():
        print("  tile period (start to next start, same CTA): mean %.0f cycles" % float(per.mean()))
