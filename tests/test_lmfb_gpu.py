"""Parity of the CUDA path (through the C ABI) against the float64 oracle, the committed
golden fixtures (incl. outputs of the live reference), and size-independent properties at
BASELINE.json's full sizes.  Tolerance: 1e-4 relative (north_star, fp32), with the metric
max|a-b| / max(|b|, rms(b)); frame counts, padding and lengths bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import lmfb_oracle as orc
import _synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4


@pytest.fixture(scope="module")
def be():
    import aas_enhancement_b200 as pkg
    from aas_enhancement_b200 import _lib
    _lib.load()                                   # fails loudly if the CUDA library is missing
    return pkg


def _fe(be, mask_mode, cmvn_mode, **kw):
    return be.LMFBFrontEnd(mask_mode=mask_mode, cmvn_mode=cmvn_mode, **kw).cuda()


def _dev(b, mask_mode, requires_grad=True):
    wave = torch.from_numpy(b["wave"]).cuda()
    lengths = torch.from_numpy(b["lengths"]).cuda()
    mr = mi = None
    if mask_mode in ("reim", "power"):
        mr = torch.from_numpy(b["mask_r"]).cuda().requires_grad_(requires_grad)
    if mask_mode == "reim":
        mi = torch.from_numpy(b["mask_i"]).cuda().requires_grad_(requires_grad)
    return wave, lengths, mr, mi


def _oracle(b, mask_mode, cmvn_mode, mel=None, window=None, grads=True):
    mel = orc.mel_filterbank() if mel is None else mel
    window = orc.hamming_window() if window is None else window
    mel = np.asarray(mel, np.float32).astype(np.float64)         # what the kernel is given
    window = np.asarray(window, np.float32).astype(np.float64)
    mr = b["mask_r"] if mask_mode != "none" else None
    mi = b["mask_i"] if mask_mode == "reim" else None
    z, fl = orc.lmfb_forward(b["wave"], b["lengths"], mr, mi, mel, window, mask_mode, cmvn_mode)
    g = None
    if grads and mask_mode != "none":
        g = orc.lmfb_grads(b["wave"], b["lengths"], mr, mi, b["grad_out"], mel, window,
                           mask_mode, cmvn_mode)
    return z, fl, g


def _run(be, b, mask_mode, cmvn_mode, fe=None, backward=True):
    fe = fe or _fe(be, mask_mode, cmvn_mode)
    wave, lengths, mr, mi = _dev(b, mask_mode)
    z, fl = fe(wave, lengths, mr, mi, tmax=b["tmax"])
    gr = gi = None
    if backward and mr is not None:
        z.backward(torch.from_numpy(b["grad_out"]).cuda())
        gr = mr.grad.cpu().numpy()
        gi = mi.grad.cpu().numpy() if mi is not None else None
    return z.detach().cpu().numpy(), fl.cpu().numpy(), gr, gi


def _check(be, b, mask_mode, cmvn_mode, tol=TOL, **fe_kw):
    fe = _fe(be, mask_mode, cmvn_mode, **fe_kw) if fe_kw else None
    z, fl, gr, gi = _run(be, b, mask_mode, cmvn_mode, fe=fe)
    mel = fe.mel_basis.cpu().numpy() if fe is not None else None
    win = fe.window.cpu().numpy() if fe is not None else None
    z_ref, fl_ref, g_ref = _oracle(b, mask_mode, cmvn_mode, mel, win)
    assert fl.dtype == np.int32 and np.array_equal(fl, fl_ref)               # bit-exact
    for i in range(len(fl)):
        assert np.all(z[i, :, fl[i]:] == 0.0)                                 # exact zero padding
    assert not np.isnan(z).any()
    assert orc.rel_err(z, z_ref) < tol
    if g_ref is not None:
        assert orc.rel_err(gr, g_ref["grad_mask_r"]) < tol
        for i in range(len(fl)):
            assert np.all(gr[i, :, fl[i]:] == 0.0)
        if gi is not None:
            assert orc.rel_err(gi, g_ref["grad_mask_i"]) < tol
    return z, gr, gi


# ------------------------------------------------------------------ oracle parity, small sizes
@pytest.mark.parametrize("mask_mode", ["reim", "power", "none"])
@pytest.mark.parametrize("cmvn_mode", ["per_bin", "global", "none"])
def test_parity_all_modes_ragged(be, mask_mode, cmvn_mode):
    b = _synth.make_batch(5, 9000, seed=41, ragged=True, tonal=(cmvn_mode == "global"))
    _check(be, b, mask_mode, cmvn_mode)


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_golden_oracle_fixtures(be, name):
    g = np.load(os.path.join(GOLD, f"oracle_lmfb_{name}.npz"))
    b = _synth.make_batch(int(g["n"]), int(g["max_len"]), seed=int(g["seed"]),
                          ragged=bool(g["ragged"]), tonal=bool(g["tonal"]))
    mm, cm = str(g["mask_mode"]), str(g["cmvn"])
    z, fl, gr, gi = _run(be, b, mm, cm)
    assert np.array_equal(fl, g["frame_lens"])
    # fixtures were made with float64 mel/window; the kernel sees their fp32 roundings
    assert orc.rel_err(z, g["z"]) < TOL
    if "grad_mask_r" in g.files:
        assert orc.rel_err(gr, g["grad_mask_r"]) < TOL
    if "grad_mask_i" in g.files:
        assert orc.rel_err(gi, g["grad_mask_i"]) < TOL


def test_against_live_reference_output(be):
    """Output and mask gradients of the reference's own BRNNmultiCH tail (model.py:186-198,
    run live when the fixture was made) vs the CUDA path on the same wave and masks."""
    g = np.load(os.path.join(GOLD, "ref_glue_on_stft.npz"))
    b = _synth.make_batch(int(g["n"]), int(g["max_len"]), seed=int(g["seed"]))
    b["mask_r"], b["mask_i"] = g["mask_real"], g["mask_imag"]
    b["grad_out"] = g["grad_out"]
    z, fl, gr, gi = _run(be, b, "reim", "none")
    assert z.shape == g["output"].shape
    assert orc.rel_err(z, g["output"]) < TOL
    assert orc.rel_err(gr, g["grad_mask_real"]) < TOL
    assert orc.rel_err(gi, g["grad_mask_imag"]) < TOL


# ------------------------------------------------------------------ BASELINE.json sizes
def test_config0_8x4s_forward(be):
    b = _synth.make_batch(8, 64000, seed=123)
    assert b["tmax"] == 401
    _check(be, b, "reim", "per_bin")


def test_config1_chime4_30x6s_fwd_bwd(be):
    b = _synth.make_batch(30, 96000, seed=123)
    assert b["tmax"] == 601
    z, gr, gi = _check(be, b, "reim", "per_bin")
    zi = z.astype(np.float64)
    assert np.abs(zi.mean(axis=2)).max() < 1e-4                     # CMVN: zero mean ...
    assert np.abs(zi.std(axis=2, ddof=1) - 1.0).max() < 1e-4       # ... unit (unbiased) std


def test_config1_ragged_tonal(be):
    b = _synth.make_batch(30, 96000, seed=7, ragged=True, tonal=True)
    _check(be, b, "reim", "per_bin")


def test_config3_paired_clean_and_noisy(be):
    """FSEGAN / minimize_DCE: masked noisy pass (fwd+bwd) + unmasked clean pass (fwd)."""
    b = _synth.make_batch(6, 48000, seed=77, ragged=True)
    _check(be, b, "reim", "per_bin")
    _check(be, b, "none", "per_bin")


def test_long_utterance_30s(be):
    b = _synth.make_batch(2, 480000, seed=5, ragged=True)
    assert b["tmax"] == 3001
    _check(be, b, "reim", "per_bin")


# ------------------------------------------------------------------ STFT as an output
@pytest.mark.parametrize("n,length", [(3, 9000), (2, 161), (40, 48000)])
def test_stft_output_matches_oracle_and_feeds_the_reference_glue(be, n, length):
    """aas_lmfb_stft: (N, 2*161, T) with the real rows first (the layout BRNNmultiCH.forward views
    as (N, 2, F, T), model.py:186-188) against the float64 oracle; pushing it through the literal
    ops of model.py:191-198 reproduces the fused forward."""
    b = _synth.make_batch(n, length, seed=length + n, ragged=True)
    fe = be.LMFBFrontEnd(mask_mode="reim", cmvn_mode="none").cuda()
    wave, lengths, mr, mi = _dev(b, "reim")
    spec, fl = fe.stft(wave, lengths)
    spec_np = spec.cpu().numpy()
    win = fe.window.cpu().numpy().astype(np.float64)
    tmax = spec.shape[2]
    ref = np.zeros((n, 2 * 161, tmax))
    for i in range(n):
        s = orc.stft_frames(b["wave"][i], int(b["lengths"][i]), win)
        ref[i, :161, :s.shape[1]] = s.real
        ref[i, 161:, :s.shape[1]] = s.imag
        assert int(fl[i]) == s.shape[1]
        assert np.all(spec_np[i, :, s.shape[1]:] == 0.0)
    assert orc.rel_err(spec_np, ref) < TOL
    st = spec.view(n, 2, 161, tmax).double().cpu()                    # model.py:186-188 (float64: cuDNN would use TF32)
    power = (st[:, 0] * mr.detach().double().cpu()) ** 2 + (st[:, 1] * mi.detach().double().cpu()) ** 2   # :191-194
    glue = torch.log1p(torch.nn.functional.conv1d(power, fe.mel_basis.double().cpu().unsqueeze(-1)))      # :196-198
    z, _ = fe(wave, lengths, mr, mi)
    assert orc.rel_err(z.detach().cpu().numpy(), glue.numpy()) < TOL


# ------------------------------------------------------------------ gradient into the waveform
@pytest.mark.parametrize("mask_mode,cmvn_mode", [("none", "per_bin"), ("reim", "per_bin"), ("power", "none"), ("reim", "global")])
@pytest.mark.parametrize("n,length", [(3, 9000), (2, 161), (48, 80000)])
def test_wave_gradient_matches_float64_autograd(be, mask_mode, cmvn_mode, n, length):
    """aas_lmfb_backward_wave through autograd: d/d wave of sum(Z * g) against the float64 oracle
    (adjoint STFT with the reflect padding folded back, overlap-add across tiles; the largest case
    is scheduled with cluster launch control), next to unchanged mask gradients."""
    if n == 48 and (mask_mode, cmvn_mode) != ("reim", "per_bin"):
        pytest.skip("one large case is enough")
    if length < 320:
        cmvn_mode = "none"          # two frames: the CMVN of two samples is +-0.707 whatever the input (no usable gradient)
    b = _synth.make_batch(n, length, seed=length + n, ragged=True)
    fe = be.LMFBFrontEnd(mask_mode=mask_mode, cmvn_mode=cmvn_mode).cuda()
    wave, lengths, mr, mi = _dev(b, mask_mode)
    wave.requires_grad_(True)
    z, fl = fe(wave, lengths, mr, mi)
    g = torch.from_numpy(b["grad_out"]).cuda()
    z.backward(g)
    mel = fe.mel_basis.cpu().numpy().astype(np.float64)
    win = fe.window.cpu().numpy().astype(np.float64)
    ref = orc.lmfb_grads(b["wave"], b["lengths"], b["mask_r"] if mr is not None else None,
                         b["mask_i"] if mi is not None else None, b["grad_out"], mel, win,
                         mask_mode=mask_mode, cmvn_mode=cmvn_mode, want_wave_grad=True)
    gw = wave.grad.cpu().numpy()
    for i in range(n):
        assert np.all(gw[i, int(b["lengths"][i]):] == 0.0)
    assert np.isfinite(gw).all()
    assert orc.rel_err(gw, ref["grad_wave"]) < TOL
    tol = TOL if length > 320 else 3e-4
    if mr is not None:
        assert orc.rel_err(mr.grad.cpu().numpy(), ref["grad_mask_r"]) < tol
    if mi is not None:
        assert orc.rel_err(mi.grad.cpu().numpy(), ref["grad_mask_i"]) < tol
    # the call without the waveform gradient gives the same mask gradients, bit for bit
    if mr is not None:
        wave2, _, mr2, mi2 = _dev(b, mask_mode)
        z2, _ = fe(wave2, lengths, mr2, mi2)
        z2.backward(g)
        assert torch.equal(mr2.grad, mr.grad)


# ------------------------------------------------------------------ properties
@pytest.mark.parametrize("seconds", [10, 17])
def test_cmvn_backward_long_rows(be, seconds):
    """Rows of 1001 / 1701 frames: the block-per-row CMVN backward variants (K = 8 / 24)."""
    b = _synth.make_batch(2, 16000 * seconds, seed=40 + seconds, ragged=True)
    _check(be, b, "reim", "per_bin")


def test_unit_masks_equal_unmasked(be):
    b = _synth.make_batch(4, 20000, seed=3, ragged=True)
    b["mask_r"][:] = 1.0
    b["mask_i"][:] = 1.0
    z0, *_ = _run(be, b, "none", "per_bin", backward=False)
    z1, *_ = _run(be, b, "reim", "per_bin", backward=False)
    z2, *_ = _run(be, b, "power", "per_bin", backward=False)
    assert np.array_equal(z0, z1) and np.array_equal(z0, z2)


def test_power_is_quadratic_in_the_wave(be):
    b = _synth.make_batch(3, 16000, seed=8, ragged=True)
    y1, *_ = _run(be, b, "reim", "none", backward=False)
    b2 = dict(b)
    b2["wave"] = b["wave"] * 2.0
    y2, *_ = _run(be, b2, "reim", "none", backward=False)
    assert orc.rel_err(np.expm1(y2.astype(np.float64)), 4.0 * np.expm1(y1.astype(np.float64))) < 1e-5


def test_deterministic_and_repeatable_backward(be):
    b = _synth.make_batch(4, 30000, seed=12, ragged=True)
    fe = _fe(be, "reim", "per_bin")
    wave, lengths, mr, mi = _dev(b, "reim")
    g = torch.from_numpy(b["grad_out"]).cuda()
    z, _ = fe(wave, lengths, mr, mi)
    z.backward(g, retain_graph=True)               # trainer_AAS.py:150 uses retain_graph=True
    g1 = mr.grad.clone()
    mr.grad = None
    mi.grad = None
    z.backward(g)
    assert torch.equal(g1, mr.grad)
    z2, _ = fe(wave, lengths, mr.detach(), mi.detach())
    assert torch.equal(z, z2)


def test_directional_derivative(be):
    """<grad, d> against a central finite difference of the forward pass along a seeded direction."""
    b = _synth.make_batch(2, 8000, seed=19)
    fe = _fe(be, "reim", "per_bin")
    wave, lengths, mr, mi = _dev(b, "reim")
    g = torch.from_numpy(b["grad_out"]).cuda()
    z, _ = fe(wave, lengths, mr, mi)
    z.backward(g)
    d = torch.from_numpy(np.random.RandomState(5).randn(*mr.shape).astype(np.float32)).cuda()
    h = 2e-2
    with torch.no_grad():
        zp, _ = fe(wave, lengths, mr + h * d, mi)
        zm, _ = fe(wave, lengths, mr - h * d, mi)
        fd = ((zp.double() - zm.double()) * g.double()).sum() / (2 * h)
        an = (mr.grad.double() * d.double()).sum()
        scale = (mr.grad.double() * d.double()).abs().sum()
    assert abs(float(fd - an)) < 2e-3 * float(scale)


# ------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("length", [1, 2, 100, 159, 160, 161, 319, 320, 321, 5119, 5120, 5121])
def test_tiny_and_boundary_lengths(be, length):
    b = _synth.make_batch(1, length, seed=length)
    cm = "per_bin" if length >= 320 else "none"       # T_i == 1 has no unbiased std (NaN, as torch)
    _check(be, b, "reim", cm)


def test_single_frame_cmvn_is_nan_like_torch(be):
    b = _synth.make_batch(1, 100, seed=1)
    z, fl, *_ = _run(be, b, "reim", "per_bin", backward=False)
    assert fl[0] == 1 and np.isnan(z[0, :, 0]).all()


def test_extra_padding_and_empty_utterance(be):
    b = _synth.make_batch(3, 4000, seed=2, lengths=[4000, 1700, 0])
    tmax = b["tmax"] + 40                               # caller's Tmax larger than any T_i
    rs = np.random.RandomState(0)
    b["mask_r"] = rs.uniform(0, 1, (3, 161, tmax)).astype(np.float32)
    b["mask_i"] = rs.uniform(0, 1, (3, 161, tmax)).astype(np.float32)
    b["grad_out"] = rs.randn(3, 40, tmax).astype(np.float32)
    b["tmax"] = tmax
    z, fl, gr, gi = _run(be, b, "reim", "per_bin")
    assert list(fl) == [26, 11, 0]
    assert np.all(z[2] == 0) and np.all(gr[2] == 0) and np.all(gi[2] == 0)
    b2 = {k: (v[:2] if isinstance(v, np.ndarray) else v) for k, v in b.items()}
    z_ref, _, g_ref = _oracle(b2, "reim", "per_bin")
    assert orc.rel_err(z[:2], z_ref) < TOL
    assert orc.rel_err(gr[:2], g_ref["grad_mask_r"]) < TOL


def test_strided_masks_and_unaligned_wave(be):
    b = _synth.make_batch(3, 7001, seed=23, ragged=True)
    fe = _fe(be, "reim", "per_bin")
    wave_store = torch.zeros(3, 7001 + 3, device="cuda")
    wave = wave_store[:, 1:7002]                          # 4-byte aligned only -> scalar staging path
    wave.copy_(torch.from_numpy(b["wave"]))
    big_r = torch.zeros(3, 161, b["tmax"] + 13, device="cuda")
    big_i = torch.zeros(3, 161, b["tmax"] + 13, device="cuda")
    mr = big_r[:, :, 5:5 + b["tmax"]]
    mi = big_i[:, :, 5:5 + b["tmax"]]
    mr.copy_(torch.from_numpy(b["mask_r"]))
    mi.copy_(torch.from_numpy(b["mask_i"]))
    mr.requires_grad_(True)
    mi.requires_grad_(True)
    z, fl = fe(wave, torch.from_numpy(b["lengths"]).cuda(), mr, mi)
    z.backward(torch.from_numpy(b["grad_out"]).cuda())
    z_ref, fl_ref, g_ref = _oracle(b, "reim", "per_bin")
    assert np.array_equal(fl.cpu().numpy(), fl_ref)
    assert orc.rel_err(z.detach().cpu().numpy(), z_ref) < TOL
    assert orc.rel_err(mr.grad.cpu().numpy(), g_ref["grad_mask_r"]) < TOL
    assert orc.rel_err(mi.grad.cpu().numpy(), g_ref["grad_mask_i"]) < TOL


@pytest.mark.parametrize("n_mels", [23, 64, 80])
def test_other_mel_bases_and_periodic_window(be, n_mels):
    b = _synth.make_batch(2, 6000, seed=n_mels, ragged=True)
    b["grad_out"] = np.random.RandomState(1).randn(2, n_mels, b["tmax"]).astype(np.float32)
    _check(be, b, "reim", "per_bin", n_mels=n_mels, window_sym=False)


def test_htk_style_custom_basis(be):
    # a caller-supplied basis (model.py:148): unnormalised HTK-like triangles over 300-7000 Hz
    freqs = np.linspace(0, 8000, 161)
    pts = 700 * (10 ** (np.linspace(2595 * np.log10(1 + 300 / 700), 2595 * np.log10(1 + 7000 / 700), 42) / 2595) - 1)
    mel = np.zeros((40, 161))
    for m in range(40):
        up = (freqs - pts[m]) / (pts[m + 1] - pts[m])
        dn = (pts[m + 2] - freqs) / (pts[m + 2] - pts[m + 1])
        mel[m] = np.maximum(0, np.minimum(up, dn))
    b = _synth.make_batch(2, 6000, seed=4, ragged=True)
    _check(be, b, "power", "global", mel_basis=mel)


def test_runs_on_a_side_stream(be):
    b = _synth.make_batch(3, 12000, seed=6, ragged=True)
    fe = _fe(be, "reim", "per_bin")
    wave, lengths, mr, mi = _dev(b, "reim")
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        z, _ = fe(wave, lengths, mr, mi)
        z.backward(torch.from_numpy(b["grad_out"]).cuda())
    s.synchronize()
    z_ref, _, g_ref = _oracle(b, "reim", "per_bin")
    assert orc.rel_err(z.detach().cpu().numpy(), z_ref) < TOL
    assert orc.rel_err(mr.grad.cpu().numpy(), g_ref["grad_mask_r"]) < TOL


def test_eps_variant(be):
    b = _synth.make_batch(2, 5000, seed=14, ragged=True)
    fe = be.LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin", eps=1e-2).cuda()
    wave, lengths, mr, mi = _dev(b, "reim")
    z, _ = fe(wave, lengths, mr, mi)
    z.backward(torch.from_numpy(b["grad_out"]).cuda())
    mel = fe.mel_basis.cpu().numpy().astype(np.float64)
    win = fe.window.cpu().numpy().astype(np.float64)
    z_ref, _ = orc.lmfb_forward(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], mel, win, eps=1e-2)
    g_ref = orc.lmfb_grads(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], b["grad_out"], mel, win, eps=1e-2)
    assert orc.rel_err(z.detach().cpu().numpy(), z_ref) < TOL
    assert orc.rel_err(mr.grad.cpu().numpy(), g_ref["grad_mask_r"]) < TOL


@pytest.mark.parametrize("warps", ["2", "3", "4", "5"])
def test_every_kernel_variant(be, warps, monkeypatch):
    """All warps-per-tile variants of K1 (tuning knob AAS_LMFB_WARPS_*) give the same answers."""
    monkeypatch.setenv("AAS_LMFB_WARPS_FWD", warps)
    monkeypatch.setenv("AAS_LMFB_WARPS_BWD", warps)
    b = _synth.make_batch(4, 11000, seed=33, ragged=True)
    for mm in ("reim", "power"):
        _check(be, b, mm, "per_bin")


# ------------------------------------------------------------------ L1Loss_mask
def test_l1loss_mask_against_live_reference(be):
    g = np.load(os.path.join(GOLD, "ref_l1loss.npz"))
    rs = np.random.RandomState(int(g["seed"]))
    n, c, tmax = 4, 40, 37
    a = torch.from_numpy(rs.randn(n, c, tmax).astype(np.float32)).cuda().requires_grad_(True)
    b = torch.from_numpy(rs.randn(n, c, tmax).astype(np.float32)).cuda().requires_grad_(True)
    mask = torch.zeros(n, 1, tmax, dtype=torch.uint8, device="cuda")
    for i, l in enumerate(g["lens"]):
        mask[i, :, int(l):] = 1
    loss, n_element = be.L1Loss_mask()(a, b, mask)
    (loss * float(g["upstream"])).backward()
    assert int(n_element) == int(g["n_element"])                       # bit-exact count (frames)
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert orc.rel_err(a.grad.cpu().numpy(), g["grad_input"]) < 1e-6
    assert orc.rel_err(b.grad.cpu().numpy(), g["grad_target"]) < 1e-6
    # deterministic
    loss2, _ = be.L1Loss_mask()(a.detach(), b.detach(), mask)
    assert float(loss2) == float(loss)
    # opt-in fix of the reference's no-op masking
    fixed, _ = be.L1Loss_mask(fix_masking=True)(a.detach(), b.detach(), mask)
    ref_fixed, _ = orc.l1loss_mask(a.detach().cpu().numpy(), b.detach().cpu().numpy(), mask.cpu().numpy(), True)
    assert abs(float(fixed) - ref_fixed) < 1e-5 * ref_fixed


def test_l1loss_on_front_end_output_full_size(be):
    """AAS G-step shape: L1 between two feature tensors of the CHiME-4-shaped batch."""
    b = _synth.make_batch(30, 96000, seed=3, ragged=True)
    fe = _fe(be, "reim", "per_bin")
    wave, lengths, mr, mi = _dev(b, "reim")
    z, fl = fe(wave, lengths, mr, mi)
    target = torch.randn_like(z)
    mask = torch.zeros(30, 1, b["tmax"], dtype=torch.uint8, device="cuda")
    for i in range(30):
        mask[i, :, int(fl[i]):] = 1
    loss, n_element = be.L1Loss_mask()(z, target, mask)
    loss.backward()
    want, want_n = orc.l1loss_mask(z.detach().cpu().numpy(), target.cpu().numpy(), mask.cpu().numpy())
    assert int(n_element) == want_n
    assert abs(loss.item() - want) < 1e-5 * want
    assert mr.grad is not None and torch.isfinite(mr.grad).all()


def test_dynamic_and_static_tile_schedules_agree():
    """A launch with more tiles than resident CTAs is scheduled with cluster launch control; the
    result must be bit-identical to the static round-robin schedule (AAS_LMFB_SCHED=static)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "check_sched.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "schedules agree" in res.stdout
