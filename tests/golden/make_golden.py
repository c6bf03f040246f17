"""Generate the committed golden fixtures under tests/golden/.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Fixtures written
  ref_glue_reim.npz   LIVE REFERENCE: BRNNmultiCH.forward (model.py:169-200) run unmodified on
                      CPU (only ``.cuda()`` in its ctor, model.py:167, is shimmed to a no-op);
                      the masks it multiplies by are captured with forward hooks; output and
                      autograd gradients w.r.t. both masks are stored.
  ref_glue_on_stft.npz  LIVE REFERENCE: the same tail fed with the oracle's STFT of a seeded wave,
                      so the CUDA path (STFT included) can be compared with the reference output.
  ref_glue_2ch.npz    LIVE REFERENCE with nCH = 2 (masks (N, 2*161, T), basis repeated over the channels).
  ref_l1loss.npz      LIVE REFERENCE: L1Loss_mask (model.py:19-31) loss, nElement and gradients.
  ref_collate.npz     LIVE REFERENCE: _collate_fn / _collate_fn_paired outputs
                      (loader_functions.py:47-105) for a seeded ragged batch.
  ref_ctc_sizes.npz   LIVE torch semantics of trainer_AAS.py:165-167 on a grid of (T, Tmax, T').
  oracle_lmfb_*.npz   float64 ORACLE outputs (parity unpinned at the STFT/CMVN boundary, see
                      oracle/lmfb_oracle.py) for seeded inputs from tests/_synth.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/Speech_enhancement_by_AAS"

from oracle import lmfb_oracle as orc   # noqa: E402
import _synth                            # noqa: E402


def make_ref_glue():
    sys.path.insert(0, REF)
    import model as ref_model            # the reference's own model.py

    torch.manual_seed(123)
    f, t, n, h = 161, 9, 2, 16
    mel = orc.mel_filterbank().astype(np.float32)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # shim for model.py:167 only
    try:
        net = ref_model.BRNNmultiCH(I=2 * f, H=h, L=2, nCH=1, mel_basis=mel)
    finally:
        torch.Tensor.cuda = orig_cuda
    net = net.float()
    captured = {}

    def hook(name):
        def fn(_m, _i, out):
            out.retain_grad()
            captured[name] = out
        return fn

    net.final_linear_real.register_forward_hook(hook("mr"))
    net.final_linear_imag.register_forward_hook(hook("mi"))
    rs = np.random.RandomState(123)
    x = torch.from_numpy(rs.randn(n, 2 * f, t).astype(np.float32) * 3.0)
    out = net(x)
    g = torch.from_numpy(rs.randn(*out.shape).astype(np.float32))
    out.backward(g)
    np.savez_compressed(
        os.path.join(HERE, "ref_glue_reim.npz"),
        stft_real=x[:, :f].numpy(), stft_imag=x[:, f:].numpy(),
        mask_real=captured["mr"].detach().numpy(), mask_imag=captured["mi"].detach().numpy(),
        mel_basis=mel, output=out.detach().numpy(), grad_out=g.numpy(),
        grad_mask_real=captured["mr"].grad.numpy(), grad_mask_imag=captured["mi"].grad.numpy())
    sys.path.remove(REF)


def make_ref_glue_on_stft():
    """The live reference tail (model.py:186-198) fed with a REAL STFT (the oracle's, of a
    seeded wave), so that the CUDA path can be compared with the reference's own output."""
    sys.path.insert(0, REF)
    import model as ref_model

    torch.manual_seed(321)
    f, h = 161, 24
    b = _synth.make_batch(2, 3200, seed=31)
    win = orc.hamming_window().astype(np.float32).astype(np.float64)
    mel = orc.mel_filterbank().astype(np.float32)
    spec = np.stack([orc.stft_frames(b["wave"][i], int(b["lengths"][i]), win) for i in range(2)])
    x = torch.from_numpy(np.concatenate([spec.real, spec.imag], axis=1).astype(np.float32))
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        net = ref_model.BRNNmultiCH(I=2 * f, H=h, L=3, nCH=1, mel_basis=mel).float()
    finally:
        torch.Tensor.cuda = orig_cuda
    # spread the masks so they are not all ~0
    with torch.no_grad():
        net.final_linear_real.bias.fill_(0.7)
        net.final_linear_imag.bias.fill_(0.6)
        net.final_linear_real.weight.mul_(3.0)
        net.final_linear_imag.weight.mul_(3.0)
    captured = {}

    def hook(name):
        def fn(_m, _i, out):
            out.retain_grad()
            captured[name] = out
        return fn

    net.final_linear_real.register_forward_hook(hook("mr"))
    net.final_linear_imag.register_forward_hook(hook("mi"))
    out = net(x)
    g = torch.from_numpy(b["grad_out"][:, :, :out.shape[2]].copy())
    out.backward(g)
    np.savez_compressed(
        os.path.join(HERE, "ref_glue_on_stft.npz"),
        seed=31, n=2, max_len=3200,
        mask_real=captured["mr"].detach().numpy(), mask_imag=captured["mi"].detach().numpy(),
        output=out.detach().numpy(), grad_out=g.numpy(),
        grad_mask_real=captured["mr"].grad.numpy(), grad_mask_imag=captured["mi"].grad.numpy())
    sys.path.remove(REF)


def make_ref_glue_2ch():
    """LIVE REFERENCE with nCH = 2: BRNNmultiCH(I = 2*2*161, nCH = 2) (model.py:148-200; the basis is
    repeated over the channels at :167, so the conv1d at :196 sums the two channels' masked powers) fed
    with the oracle's STFT of a seeded two-channel wave."""
    sys.path.insert(0, REF)
    import model as ref_model

    torch.manual_seed(77)
    f, h, n_ch = 161, 20, 2
    b0 = _synth.make_batch(2, 2900, seed=41, ragged=True)
    b1 = _synth.make_batch(2, 2900, seed=42, lengths=b0["lengths"])
    win = orc.hamming_window().astype(np.float32).astype(np.float64)
    mel = orc.mel_filterbank().astype(np.float32)
    tmax = b0["tmax"]
    re = np.zeros((2, n_ch * f, tmax)); im = np.zeros((2, n_ch * f, tmax))
    for i in range(2):
        for c, b in enumerate((b0, b1)):
            sp = orc.stft_frames(b["wave"][i], int(b["lengths"][i]), win)
            re[i, c * f:(c + 1) * f, :sp.shape[1]] = sp.real
            im[i, c * f:(c + 1) * f, :sp.shape[1]] = sp.imag
    x = torch.from_numpy(np.concatenate([re, im], axis=1).astype(np.float32))     # (N, nCH*F*2, T), model.py:170
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        net = ref_model.BRNNmultiCH(I=2 * n_ch * f, H=h, L=2, nCH=n_ch, mel_basis=mel).float()
    finally:
        torch.Tensor.cuda = orig_cuda
    with torch.no_grad():
        net.final_linear_real.bias.fill_(0.7)
        net.final_linear_imag.bias.fill_(0.6)
        net.final_linear_real.weight.mul_(3.0)
        net.final_linear_imag.weight.mul_(3.0)
    captured = {}

    def hook(name):
        def fn(_m, _i, out):
            out.retain_grad()
            captured[name] = out
        return fn

    net.final_linear_real.register_forward_hook(hook("mr"))
    net.final_linear_imag.register_forward_hook(hook("mi"))
    out = net(x)
    g = torch.from_numpy(b0["grad_out"][:, :, :out.shape[2]].copy())
    out.backward(g)
    np.savez_compressed(
        os.path.join(HERE, "ref_glue_2ch.npz"),
        seeds=np.asarray([41, 42]), n=2, max_len=2900,
        mask_real=captured["mr"].detach().numpy(), mask_imag=captured["mi"].detach().numpy(),
        output=out.detach().numpy(), grad_out=g.numpy(),
        grad_mask_real=captured["mr"].grad.numpy(), grad_mask_imag=captured["mi"].grad.numpy())
    sys.path.remove(REF)


def make_ref_l1loss():
    """LIVE REFERENCE: L1Loss_mask.forward (model.py:19-31) + autograd on a padded, length-sorted batch."""
    sys.path.insert(0, REF)
    import model as ref_model
    rs = np.random.RandomState(99)
    n, c, tmax = 4, 40, 37
    lens = [37, 30, 22, 9]
    a = torch.from_numpy(rs.randn(n, c, tmax).astype(np.float32)).requires_grad_(True)
    b = torch.from_numpy(rs.randn(n, c, tmax).astype(np.float32)).requires_grad_(True)
    mask = torch.zeros(n, 1, tmax, dtype=torch.uint8)
    for i, l in enumerate(lens):
        mask[i, :, l:] = 1
    # torch >= 1.2 rejects the reference's ByteTensor mask inside masked_fill (model.py:29), so the
    # reference is run with the same mask as bool; every other line executes unmodified
    loss, n_element = ref_model.L1Loss_mask()(a, b, mask.bool())
    (loss * 3.0).backward()
    np.savez_compressed(os.path.join(HERE, "ref_l1loss.npz"), seed=99, lens=np.asarray(lens),
                        loss=loss.detach().numpy(), n_element=int(n_element),
                        grad_input=a.grad.numpy(), grad_target=b.grad.numpy(), upstream=3.0)
    sys.path.remove(REF)


def make_ref_collate():
    sys.path.insert(0, REF)
    import loader_functions as lf

    rs = np.random.RandomState(7)
    lens = [37, 52, 52, 11, 29, 52, 1]
    batch, paired = [], []
    for i, t in enumerate(lens):
        feat = torch.from_numpy(rs.randn(40, t).astype(np.float32))
        clean = torch.from_numpy(rs.randn(40, t).astype(np.float32))
        txt = [int(v) for v in rs.randint(1, 29, size=3 + i)]
        batch.append((feat, txt))
        paired.append((feat, txt, clean))
    inputs, targets, pct, tsz, mask = lf._collate_fn(list(batch))
    pi, po, pm, pt, ppct, ptsz = lf._collate_fn_paired(list(paired))
    np.savez_compressed(
        os.path.join(HERE, "ref_collate.npz"),
        lens=np.asarray(lens), seed=7,
        inputs=inputs.numpy(), targets=targets.numpy(), pct=pct.numpy(), tsz=tsz.numpy(),
        mask=mask.numpy(),
        p_inputs=pi.numpy(), p_outputs=po.numpy(), p_mask=pm.numpy(), p_targets=pt.numpy(),
        p_pct=ppct.numpy(), p_tsz=ptsz.numpy())
    sys.path.remove(REF)


def make_ref_ctc_sizes():
    rows = []
    for tmax in (101, 200, 401, 435, 601, 1234, 3001):
        t_out = orc.conv_out_frames(tmax)
        for t in sorted(set([1, 2, 57, 100, 135, tmax // 3, tmax // 2, tmax - 1, tmax])):
            if t > tmax:
                continue
            pct = torch.FloatTensor(1)
            pct[0] = t / float(tmax)                          # loader_functions.py:57
            sizes = pct.mul_(int(t_out)).int()                # trainer_AAS.py:165-167
            rows.append((t, tmax, t_out, int(sizes[0])))
    np.savez_compressed(os.path.join(HERE, "ref_ctc_sizes.npz"), rows=np.asarray(rows, dtype=np.int64))


def make_oracle_cases():
    cases = {
        "a": dict(n=3, max_len=4000, ragged=True, seed=123, mask_mode="reim", cmvn="per_bin"),
        "b": dict(n=2, max_len=2777, ragged=True, seed=5, tonal=True, mask_mode="power", cmvn="global"),
        "c": dict(n=2, max_len=1600, ragged=False, seed=9, mask_mode="none", cmvn="none"),
        "d": dict(n=2, max_len=3333, ragged=True, seed=11, tonal=True, mask_mode="reim", cmvn="none"),
    }
    for name, c in cases.items():
        b = _synth.make_batch(c["n"], c["max_len"], seed=c["seed"], ragged=c["ragged"],
                              tonal=c.get("tonal", False))
        mr = b["mask_r"] if c["mask_mode"] in ("reim", "power") else None
        mi = b["mask_i"] if c["mask_mode"] == "reim" else None
        z, fl = orc.lmfb_forward(b["wave"], b["lengths"], mr, mi, mask_mode=c["mask_mode"],
                                 cmvn_mode=c["cmvn"])
        out = dict(z=z, frame_lens=fl, mask_mode=c["mask_mode"], cmvn=c["cmvn"],
                   n=c["n"], max_len=c["max_len"], ragged=c["ragged"], seed=c["seed"],
                   tonal=c.get("tonal", False))
        if c["mask_mode"] != "none":
            g = orc.lmfb_grads(b["wave"], b["lengths"], mr, mi, b["grad_out"],
                               mask_mode=c["mask_mode"], cmvn_mode=c["cmvn"])
            out["grad_mask_r"] = g["grad_mask_r"]
            if "grad_mask_i" in g:
                out["grad_mask_i"] = g["grad_mask_i"]
        np.savez_compressed(os.path.join(HERE, f"oracle_lmfb_{name}.npz"), **out)


if __name__ == "__main__":
    make_ref_glue()
    make_ref_glue_on_stft()
    make_ref_glue_2ch()
    make_ref_l1loss()
    make_ref_collate()
    make_ref_ctc_sizes()
    make_oracle_cases()
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))
