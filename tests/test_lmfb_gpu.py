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


@pytest.mark.parametrize("warps", [4, 5, 6, 8])
def test_every_kernel_variant(be, warps):
    """All warps-per-tile variants of K1 (aas_lmfb_plan_set_tuning) give the same answers."""
    b = _synth.make_batch(4, 11000, seed=33, ragged=True)
    for mm in ("reim", "power"):
        fe = _fe(be, mm, "per_bin").set_tuning(warps, warps)
        z, fl, gr, gi = _run(be, b, mm, "per_bin", fe=fe)
        z_ref, fl_ref, g_ref = _oracle(b, mm, "per_bin")
        assert np.array_equal(fl, fl_ref)
        assert orc.rel_err(z, z_ref) < TOL
        assert orc.rel_err(gr, g_ref["grad_mask_r"]) < TOL
        if gi is not None:
            assert orc.rel_err(gi, g_ref["grad_mask_i"]) < TOL


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
    result must be bit-identical to the static round-robin schedule (`set_tuning(static_schedule=True)`)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "check_sched.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "schedules agree" in res.stdout


# ------------------------------------------------------------------ parity AT the sizes the roofline claims are made on
def _big_batch(n, seconds, seed, ragged=True):
    """Seeded batch generated on the device (the oracle only ever sees the utterances it checks)."""
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    samples = int(seconds * 16000)
    tmax = 1 + samples // 160
    wave = (0.1 * torch.randn(n, samples, generator=gen, device="cuda")).clamp_(-1, 1)
    if ragged:
        rs = np.random.RandomState(seed)
        lengths = (rs.uniform(0.5, 1.0, size=n) * samples).astype(np.int64)
        lengths[0] = samples
        lengths = np.sort(lengths)[::-1].copy()
    else:
        lengths = np.full(n, samples, dtype=np.int64)
    lens = torch.from_numpy(lengths.astype(np.int32)).cuda()
    wave *= (torch.arange(samples, device="cuda")[None, :] < lens[:, None])       # zero padded like the collate
    mr = torch.rand(n, 161, tmax, generator=gen, device="cuda")
    mi = torch.rand(n, 161, tmax, generator=gen, device="cuda")
    g = torch.randn(n, 40, tmax, generator=gen, device="cuda")
    return wave, lens, mr, mi, g, tmax, lengths


def _oracle_rows(wave, lens, mr, mi, g, rows, mask_mode="reim", cmvn_mode="per_bin"):
    """float64 oracle (features + mask gradients) of the chosen utterances only."""
    idx = torch.as_tensor(rows, device="cuda")
    b = dict(wave=wave[idx].cpu().numpy(), lengths=lens[idx].cpu().numpy(),
             mask_r=mr[idx].cpu().numpy() if mr is not None else None,
             mask_i=mi[idx].cpu().numpy() if mi is not None else None,
             grad_out=g[idx].cpu().numpy())
    return _oracle(b, mask_mode, cmvn_mode, grads=mr is not None)


@pytest.mark.parametrize("ragged", [False, True])
def test_parity_at_sweep_256x10s_dynamic_schedule(be, ragged):
    """256 x 10 s = 8,192 tiles: far more than the resident CTAs, so both K1 launches run under the
    cluster-launch-control schedule.  Compared with the ORACLE on a seeded subset of utterances
    (every tile of those utterances), plus the size-independent properties on the whole batch."""
    wave, lens, mr, mi, g, tmax, lengths = _big_batch(256, 10.0, seed=2560 + int(ragged), ragged=ragged)
    assert tmax == 1001
    fe = _fe(be, "reim", "per_bin")
    mr.requires_grad_(True)
    mi.requires_grad_(True)
    z, fl = fe(wave, lens, mr, mi)
    z.backward(g)
    fl_ref = np.minimum(1 + lengths // 160, tmax).astype(np.int32)
    assert np.array_equal(fl.cpu().numpy(), fl_ref)                                   # bit-exact frame counts
    t = torch.arange(tmax, device="cuda")[None, None, :]
    pad = t >= fl[:, None, None]
    assert not bool((z.detach() * pad).abs().sum() != 0)                              # exact zero padding ...
    assert not bool((mr.grad * pad).abs().sum() != 0) and not bool((mi.grad * pad).abs().sum() != 0)
    assert bool(torch.isfinite(z).all()) and bool(torch.isfinite(mr.grad).all())
    zd = z.detach().double()
    cnt = fl[:, None].double()
    mean = zd.sum(dim=2) / cnt                                                        # ... and CMVN over each utterance's own frames
    assert float(mean.abs().max()) < 1e-4
    var = ((zd - mean[:, :, None]) ** 2 * (~pad)).sum(dim=2) / (cnt - 1)
    assert float((var.sqrt() - 1).abs().max()) < 1e-4
    rows = sorted(set([0, 255] + list(np.random.RandomState(7).choice(256, 3, replace=False))))
    z_ref, _, g_ref = _oracle_rows(wave, lens, mr.detach(), mi.detach(), g, rows)
    idx = torch.as_tensor(rows, device="cuda")
    assert orc.rel_err(z.detach()[idx].cpu().numpy(), z_ref) < TOL
    assert orc.rel_err(mr.grad[idx].cpu().numpy(), g_ref["grad_mask_r"]) < TOL
    assert orc.rel_err(mi.grad[idx].cpu().numpy(), g_ref["grad_mask_i"]) < TOL


def test_parity_paired_30x6s_both_passes(be):
    """BASELINE.json configs[3] at full size: masked noisy pass (forward + backward) and unmasked
    clean pass (forward) of the same 30 utterances, every element against the oracle."""
    b = _synth.make_batch(30, 96000, seed=303, ragged=True)
    _check(be, b, "reim", "per_bin")
    clean = _synth.make_batch(30, 96000, seed=304, lengths=b["lengths"])
    zc, *_ = _check(be, clean, "none", "per_bin")
    assert zc.shape == (30, 40, 601)


def test_parity_at_sweep_512x30s(be):
    """The largest sweep point: 512 x 30 s (1.5 M frames, 48,128 tiles).  Frame counts, zero padding
    and lengths bit-exact on the whole batch; features and gradients of a row subset vs the oracle."""
    wave, lens, mr, mi, g, tmax, lengths = _big_batch(512, 30.0, seed=51230, ragged=True)
    assert tmax == 3001
    fe = _fe(be, "reim", "per_bin")
    mr.requires_grad_(True)
    mi.requires_grad_(True)
    z, fl = fe(wave, lens, mr, mi)
    z.backward(g)
    fl_ref = np.minimum(1 + lengths // 160, tmax).astype(np.int32)
    assert fl.dtype == torch.int32 and np.array_equal(fl.cpu().numpy(), fl_ref)
    t = torch.arange(tmax, device="cuda")[None, None, :]
    pad = t >= fl[:, None, None]
    assert not bool((z.detach() * pad).abs().sum() != 0)
    assert not bool((mr.grad * pad).abs().sum() != 0) and not bool((mi.grad * pad).abs().sum() != 0)
    rows = [0, 257, 511]
    z_ref, _, g_ref = _oracle_rows(wave, lens, mr.detach(), mi.detach(), g, rows)
    idx = torch.as_tensor(rows, device="cuda")
    assert orc.rel_err(z.detach()[idx].cpu().numpy(), z_ref) < TOL
    assert orc.rel_err(mr.grad[idx].cpu().numpy(), g_ref["grad_mask_r"]) < TOL
    assert orc.rel_err(mi.grad[idx].cpu().numpy(), g_ref["grad_mask_i"]) < TOL


# ------------------------------------------------------------------ host-side robustness (advisor findings, round 1)
def test_wave_gradient_of_a_strided_wave_view(be):
    """wave = big[:, :L] (row stride > L) with requires_grad: the gradient buffer must take the
    wave's row stride (zeros_like would not), and nothing may be written outside it."""
    b = _synth.make_batch(3, 5000, seed=91, ragged=True)
    store = torch.zeros(3, 5000 + 64, device="cuda")
    wave = store[:, :5000]
    wave.copy_(torch.from_numpy(b["wave"]))
    wave = wave.detach().requires_grad_(True)
    assert wave.stride(0) == 5064
    fe = _fe(be, "reim", "per_bin")
    _, lengths, mr, mi = _dev(b, "reim")
    z, _ = fe(wave, lengths, mr, mi)
    z.backward(torch.from_numpy(b["grad_out"]).cuda())
    mel = fe.mel_basis.cpu().numpy().astype(np.float64)
    win = fe.window.cpu().numpy().astype(np.float64)
    ref = orc.lmfb_grads(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], b["grad_out"], mel, win,
                         want_wave_grad=True)
    assert wave.grad.shape == (3, 5000)
    assert orc.rel_err(wave.grad.cpu().numpy(), ref["grad_wave"]) < TOL
    assert orc.rel_err(mr.grad.cpu().numpy(), ref["grad_mask_r"]) < TOL


def test_mask_expanded_over_the_batch(be):
    """A (1, 161, T) mask expanded to N utterances (batch stride 0): gradients of the utterances must
    not race on the same words; autograd's expand-backward then SUMS them."""
    b = _synth.make_batch(3, 4000, seed=92)
    b["mask_r"][:] = b["mask_r"][0]
    b["mask_i"][:] = b["mask_i"][0]
    fe = _fe(be, "reim", "per_bin")
    wave, lengths, _, _ = _dev(b, "reim")
    m1r = torch.from_numpy(b["mask_r"][:1]).cuda().requires_grad_(True)
    m1i = torch.from_numpy(b["mask_i"][:1]).cuda().requires_grad_(True)
    z, _ = fe(wave, lengths, m1r.expand(3, -1, -1), m1i.expand(3, -1, -1))
    z.backward(torch.from_numpy(b["grad_out"]).cuda())
    _, _, g_ref = _oracle(b, "reim", "per_bin")
    assert orc.rel_err(m1r.grad.cpu().numpy(), g_ref["grad_mask_r"].sum(axis=0, keepdims=True)) < TOL
    assert orc.rel_err(m1i.grad.cpu().numpy(), g_ref["grad_mask_i"].sum(axis=0, keepdims=True)) < TOL


def test_plan_follows_the_mel_basis_buffer(be):
    """load_state_dict with another basis, deepcopy and torch.save of the module: the native plan is
    rebuilt from the buffer, never stale, never pickled."""
    import copy
    import io
    b = _synth.make_batch(2, 5000, seed=93, ragged=True)
    b["grad_out"] = np.random.RandomState(2).randn(2, 40, b["tmax"]).astype(np.float32)
    fe = _fe(be, "reim", "none")
    other = be.slaney_mel_basis(fmin=100.0, fmax=7000.0)
    donor = be.LMFBFrontEnd(mel_basis=other, mask_mode="reim", cmvn_mode="none")
    fe.load_state_dict(donor.state_dict())
    wave, lengths, mr, mi = _dev(b, "reim", requires_grad=False)
    z, _ = fe(wave, lengths, mr, mi)
    z_ref, _, _ = _oracle(b, "reim", "none", mel=other, grads=False)
    assert orc.rel_err(z.cpu().numpy(), z_ref) < TOL
    fe2 = copy.deepcopy(fe)
    z2, _ = fe2(wave, lengths, mr, mi)
    assert torch.equal(z, z2)
    buf = io.BytesIO()
    torch.save(fe, buf)
    buf.seek(0)
    fe3 = torch.load(buf, weights_only=False).cuda()
    z3, _ = fe3(wave, lengths, mr, mi)
    assert torch.equal(z, z3)


def test_wave_loader_to_device_front_end(be, tmp_path):
    """SURVEY 8(f) rank 3 on the GPU: manifest -> WaveDataLoader (one worker process, batches pinned in
    the main process) -> to_device on a side stream -> LMFBFrontEnd, against the oracle; int16 files
    (bytes on the wire halved) included."""
    from aas_enhancement_b200 import WaveDataLoader, save_wave, to_device
    rs = np.random.RandomState(5)
    labels = "_'ABCDEFGHIJKLMNOPQRSTUVWXYZ "
    lines, waves = [], {}
    for i in range(6):
        li = 160 * (40 + 9 * i) + 3 * i
        w = (0.3 * rs.randn(li)).clip(-1, 1).astype(np.float32)
        if i % 2:
            w = (np.round(w * 32768).clip(-32768, 32767) / 32768).astype(np.float32)   # exactly representable in int16
        wp, tp = tmp_path / f"u{i}.pt", tmp_path / f"u{i}.txt"
        save_wave(str(wp), w, int16=bool(i % 2))
        tp.write_text("HELLO\n", encoding="utf8")
        lines.append(f"{wp},{tp}")
        waves[li] = w
    man = tmp_path / "m.csv"
    man.write_text("\n".join(lines) + "\n")
    dl = WaveDataLoader(batch_size=6, tr_ny_manifest=str(man), labels=labels, num_workers=1)
    batch = dl.next("ny", "train")
    assert batch[0].is_pinned()
    side = torch.cuda.Stream()
    inputs, targets, pct, sizes, mask, lengths = to_device(batch, stream=side)
    torch.cuda.current_stream().wait_stream(side)
    fe = _fe(be, "none", "per_bin")
    z, fl = fe(inputs, lengths)
    ls = batch[5].numpy()
    assert list(ls) == sorted(ls, reverse=True)
    b = dict(wave=batch[0].numpy(), lengths=ls)
    for i, li in enumerate(ls):
        assert np.array_equal(b["wave"][i, :li], waves[int(li)])
    z_ref, fl_ref = orc.lmfb_forward(b["wave"], ls, None, None, fe.mel_basis.cpu().numpy().astype(np.float64),
                                     fe.window.cpu().numpy().astype(np.float64), "none", "per_bin")
    assert np.array_equal(fl.cpu().numpy(), fl_ref)
    assert orc.rel_err(z.cpu().numpy(), z_ref) < TOL
    assert np.array_equal((mask[:, 0].cpu().numpy() == 1), np.arange(z.shape[2])[None, :] >= fl_ref[:, None])


# ------------------------------------------------------------------ any mel basis, any number of channels (model.py:167, :196)
@pytest.mark.parametrize("kind", ["ones", "dense", "permuted", "gappy"])
@pytest.mark.parametrize("mask_mode,cmvn_mode", [("reim", "per_bin"), ("power", "none")])
def test_any_mel_basis(be, kind, mask_mode, cmvn_mode):
    """The reference applies ANY (40, F) matrix with a k=1 conv1d: all-ones, dense random, re-ordered
    filters and rows with holes / an empty filter run on the generic paths and match the oracle."""
    rs = np.random.RandomState(11)
    if kind == "ones":
        mel = np.ones((40, 161)) * 0.01
    elif kind == "dense":
        mel = rs.rand(40, 161) * 0.02
    elif kind == "permuted":
        mel = orc.mel_filterbank()[rs.permutation(40)]
    else:
        full = orc.mel_filterbank()
        keep = rs.rand(40, 161) > 0.3
        keep[np.arange(40), full.argmax(axis=1)] = True      # (an all-zero filter has zero variance: CMVN is 0/0, as in torch)
        mel = full * keep
        mel[7] = 0.0
        mel[7, 30] = 0.01                              # a filter far from its neighbours' bins
    if cmvn_mode == "per_bin" and kind == "ones":
        mel = mel * np.linspace(0.5, 1.5, 40)[:, None]   # (identical rows are fine; just make them distinguishable)
    b = _synth.make_batch(3, 7000, seed=5, ragged=True)
    _check(be, b, mask_mode, cmvn_mode, mel_basis=mel)


def test_mel_plan_of_ones_like_the_verdict_asks(be):
    assert be.MelPlan(np.ones((40, 161))).handle


@pytest.mark.parametrize("n_ch,mask_mode", [(2, "reim"), (3, "power"), (2, "none")])
def test_multi_channel_matches_oracle(be, n_ch, mask_mode):
    rs = np.random.RandomState(n_ch)
    base = _synth.make_batch(3, 9000, seed=60 + n_ch, ragged=True)
    n, tmax, lmax = 3, base["tmax"], base["wave"].shape[1]
    wave = np.zeros((n, n_ch, lmax), np.float32)
    for i in range(n):
        li = int(base["lengths"][i])
        wave[i, :, :li] = (0.1 * rs.randn(n_ch, li)).astype(np.float32)
    mr = rs.uniform(0, 1, (n, n_ch * 161, tmax)).astype(np.float32)
    mi = rs.uniform(0, 1, (n, n_ch * 161, tmax)).astype(np.float32)
    fe = _fe(be, mask_mode, "per_bin")
    w_d = torch.from_numpy(wave).cuda()
    l_d = torch.from_numpy(base["lengths"]).cuda()
    mr_d = torch.from_numpy(mr).cuda().requires_grad_(True) if mask_mode != "none" else None
    mi_d = torch.from_numpy(mi).cuda().requires_grad_(True) if mask_mode == "reim" else None
    z, fl = fe(w_d, l_d, mr_d, mi_d)
    mel = fe.mel_basis.cpu().numpy().astype(np.float64)
    win = fe.window.cpu().numpy().astype(np.float64)
    z_ref, fl_ref = orc.lmfb_forward(wave, base["lengths"], mr if mr_d is not None else None,
                                     mi if mi_d is not None else None, mel, win, mask_mode, "per_bin")
    assert np.array_equal(fl.cpu().numpy(), fl_ref)
    assert orc.rel_err(z.detach().cpu().numpy(), z_ref) < TOL
    if mr_d is None:
        return
    z.backward(torch.from_numpy(base["grad_out"]).cuda())
    g_ref = orc.lmfb_grads(wave, base["lengths"], mr, mi if mi_d is not None else None, base["grad_out"], mel, win,
                           mask_mode, "per_bin")
    assert orc.rel_err(mr_d.grad.cpu().numpy(), g_ref["grad_mask_r"]) < TOL
    if mi_d is not None:
        assert orc.rel_err(mi_d.grad.cpu().numpy(), g_ref["grad_mask_i"]) < TOL


def test_two_channels_against_live_reference(be):
    """Output and mask gradients of the reference's own BRNNmultiCH(nCH=2) tail (model.py:167, :186-198,
    run live when the fixture was made) vs the CUDA path on the same two-channel wave and masks."""
    g = np.load(os.path.join(GOLD, "ref_glue_2ch.npz"))
    b0 = _synth.make_batch(2, 2900, seed=int(g["seeds"][0]), ragged=True)
    b1 = _synth.make_batch(2, 2900, seed=int(g["seeds"][1]), lengths=b0["lengths"])
    wave = torch.from_numpy(np.stack([b0["wave"], b1["wave"]], axis=1)).cuda()       # (N, 2, L)
    mr = torch.from_numpy(g["mask_real"]).cuda().requires_grad_(True)
    mi = torch.from_numpy(g["mask_imag"]).cuda().requires_grad_(True)
    fe = _fe(be, "reim", "none")
    z, fl = fe(wave, torch.from_numpy(b0["lengths"]).cuda(), mr, mi)
    z.backward(torch.from_numpy(g["grad_out"]).cuda())
    assert z.shape == g["output"].shape
    assert orc.rel_err(z.detach().cpu().numpy(), g["output"]) < TOL
    assert orc.rel_err(mr.grad.cpu().numpy(), g["grad_mask_real"]) < TOL
    assert orc.rel_err(mi.grad.cpu().numpy(), g["grad_mask_imag"]) < TOL


def test_oversized_lengths_are_clamped_not_read(be):
    """lengths[i] > wave.shape[1] (a caller bug) must not read past the rows: the library clamps to the
    readable length it is given."""
    b = _synth.make_batch(2, 4000, seed=71)
    fe = _fe(be, "none", "none")
    wave = torch.from_numpy(b["wave"]).cuda()
    big = torch.tensor([4000 + 5000, 4000], dtype=torch.int32, device="cuda")
    z, fl = fe(wave, big)
    z2, fl2 = fe(wave, torch.from_numpy(b["lengths"]).cuda())
    assert torch.equal(z, z2) and torch.equal(fl, fl2)


def test_int16_wave_is_converted_in_the_kernel(be):
    """int16 PCM on the wire and in HBM; the kernel converts (x / 32768).  Bit-identical
    to the float path fed with the same values converted on the host; boundary tiles, ragged lengths, a
    row stride that is not a multiple of 8 samples (no bulk copies) and a CLC-scheduled launch included."""
    rs = np.random.RandomState(3)
    for n, lmax, pad in ((4, 20000, 0), (3, 7003, 3), (60, 80000, 0)):
        base = _synth.make_batch(n, lmax, seed=80 + n, ragged=True)
        pcm = np.round(base["wave"] * 32768.0).clip(-32768, 32767).astype(np.int16)
        store = torch.zeros(n, lmax + pad, dtype=torch.int16, device="cuda")
        w16 = store[:, :lmax]
        w16.copy_(torch.from_numpy(pcm))
        wf = torch.from_numpy(pcm.astype(np.float32) / 32768.0).cuda()
        lengths = torch.from_numpy(base["lengths"]).cuda()
        # (the int16 kernels exist in the five-warp shape only; pin the float path to it: the mel sums are
        # split between the warps of a tile, so bit-identity holds per shape)
        fe = _fe(be, "reim", "per_bin").set_tuning(5, 5)
        outs = []
        for wave in (w16, wf):
            mr = torch.from_numpy(base["mask_r"]).cuda().requires_grad_(True)
            mi = torch.from_numpy(base["mask_i"]).cuda().requires_grad_(True)
            z, fl = fe(wave, lengths, mr, mi)
            z.backward(torch.from_numpy(base["grad_out"]).cuda())
            outs.append((z.detach(), fl, mr.grad, mi.grad))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
        if n == 4:
            b2 = dict(base)
            b2["wave"] = pcm.astype(np.float32) / 32768.0
            z_ref, fl_ref, g_ref = _oracle(b2, "reim", "per_bin")
            assert orc.rel_err(outs[0][0].cpu().numpy(), z_ref) < TOL
            assert orc.rel_err(outs[0][2].cpu().numpy(), g_ref["grad_mask_r"]) < TOL
        zc, _ = _fe(be, "none", "per_bin").set_tuning(5, 5)(w16, lengths)
        zf, _ = _fe(be, "none", "per_bin").set_tuning(5, 5)(wf, lengths)
        assert torch.equal(zc, zf)


def test_utterance_longer_than_the_register_resident_cmvn(be):
    """34 s = 3,401 frames: beyond the block-per-row register-resident CMVN kernels (3,072), i.e. the
    three-pass fallback, forward and backward."""
    b = _synth.make_batch(1, 16000 * 34, seed=34)
    assert b["tmax"] == 3401
    _check(be, b, "reim", "per_bin")


@pytest.mark.parametrize("seconds", [3, 17, 34])
def test_l1_epilogue_of_the_cmvn_kernel(be, seconds):
    """SURVEY 8(f) rank 1 as written: L1Loss_mask (model.py:19-31) as an epilogue of the CMVN kernel.  The
    per-row sums |Z - target| are formed while Z is in registers (warp-per-row, block-per-row and the
    three-pass kernel: 301 / 1,701 / 3,401 frames); loss, nElement and the gradients equal the stand-alone
    loss on the same tensors (value to float32 summation-order accuracy) and the oracle."""
    b = _synth.make_batch(3, 16000 * seconds, seed=50 + seconds, ragged=True)
    fe = _fe(be, "reim", "per_bin")
    wave, lengths, mr, mi = _dev(b, "reim")
    rs = np.random.RandomState(seconds)
    target = torch.from_numpy(rs.randn(3, 40, b["tmax"]).astype(np.float32)).cuda()
    z, fl, rows = fe(wave, lengths, mr, mi, l1_target=target)
    mask = torch.zeros(3, 1, b["tmax"], dtype=torch.uint8, device="cuda")
    for i in range(3):
        mask[i, :, int(fl[i]):] = 1
    loss_f, n_f = be.L1Loss_mask()(z, target, mask, rows=rows)
    loss_f.backward()
    g_fused = mr.grad.clone()
    mr.grad = None
    mi.grad = None
    z2, _ = fe(wave, lengths, mr, mi)
    assert torch.equal(z, z2)                                            # the epilogue does not change Z
    loss_s, n_s = be.L1Loss_mask()(z2, target, mask)
    loss_s.backward()
    assert int(n_f) == int(n_s)
    assert abs(float(loss_f) - float(loss_s)) < 2e-6 * abs(float(loss_s))
    assert torch.equal(g_fused, mr.grad)                                 # same gradient kernels, same scale ...
    want, want_n = orc.l1loss_mask(z.detach().cpu().numpy(), target.cpu().numpy(), mask.cpu().numpy())
    assert int(n_f) == want_n and abs(float(loss_f) - want) < 1e-5 * want
    # per-row sums against numpy
    ref_rows = np.abs(z.detach().cpu().numpy().astype(np.float64) - target.cpu().numpy()).sum(axis=2)
    assert orc.rel_err(rows.cpu().numpy(), ref_rows) < 1e-5
