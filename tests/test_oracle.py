"""The oracle against the reference's live outputs (committed fixtures) and against
independent library implementations.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import lmfb_oracle as orc
import _synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference/Speech_enhancement_by_AAS"


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def test_glue_matches_live_reference_fixture():
    g = _load("ref_glue_reim.npz")
    out = orc.glue_reim(g["stft_real"].astype(np.float64), g["stft_imag"].astype(np.float64),
                        g["mask_real"].astype(np.float64), g["mask_imag"].astype(np.float64),
                        g["mel_basis"].astype(np.float64))
    assert out.shape == g["output"].shape
    assert orc.rel_err(g["output"], out) < 2e-6        # reference ran in fp32


def test_glue_gradients_match_live_reference_fixture():
    g = _load("ref_glue_reim.npz")
    re, im = g["stft_real"].astype(np.float64), g["stft_imag"].astype(np.float64)
    mr, mi = g["mask_real"].astype(np.float64), g["mask_imag"].astype(np.float64)
    mel = g["mel_basis"].astype(np.float64)
    e = np.einsum("mf,nft->nmt", mel, (re * mr) ** 2 + (im * mi) ** 2)
    d_e = g["grad_out"] / (1.0 + e)
    d_p = np.einsum("mf,nmt->nft", mel, d_e)
    assert orc.rel_err(g["grad_mask_real"], 2 * mr * re * re * d_p) < 1e-5
    assert orc.rel_err(g["grad_mask_imag"], 2 * mi * im * im * d_p) < 1e-5


def _rebuild_collate_batch():
    rs = np.random.RandomState(7)
    lens = [37, 52, 52, 11, 29, 52, 1]
    batch, paired = [], []
    for i, t in enumerate(lens):
        feat = rs.randn(40, t).astype(np.float32)
        clean = rs.randn(40, t).astype(np.float32)
        txt = [int(v) for v in rs.randint(1, 29, size=3 + i)]
        batch.append((feat, txt))
        paired.append((feat, txt, clean))
    return batch, paired


def test_collate_bit_exact_vs_reference_fixture():
    g = _load("ref_collate.npz")
    batch, paired = _rebuild_collate_batch()
    inputs, targets, pct, tsz, mask = orc.collate(batch)
    for a, b in ((inputs, g["inputs"]), (targets, g["targets"]), (pct, g["pct"]),
                 (tsz, g["tsz"]), (mask, g["mask"])):
        assert a.dtype == b.dtype and a.shape == b.shape
        assert np.array_equal(a, b)
    pi, po, pm, pt, ppct, ptsz = orc.collate_paired(paired)
    for a, b in ((pi, g["p_inputs"]), (po, g["p_outputs"]), (pm, g["p_mask"]),
                 (pt, g["p_targets"]), (ppct, g["p_pct"]), (ptsz, g["p_tsz"])):
        assert a.dtype == b.dtype and a.shape == b.shape
        assert np.array_equal(a, b)


def test_ctc_sizes_bit_exact_vs_reference_fixture():
    rows = _load("ref_ctc_sizes.npz")["rows"]
    assert len(rows) > 40
    for t, tmax, t_out, want in rows:
        assert orc.conv_out_frames(int(tmax)) == t_out
        pct = np.float32(t / float(tmax))
        assert int(orc.ctc_sizes(np.asarray([pct]), int(t_out))[0]) == want
    # the float32 path differs from exact integer arithmetic
    pct = np.float32(135 / float(435))
    assert int(orc.ctc_sizes(np.asarray([pct]), 203)[0]) == 62
    assert (135 * 203) // 435 == 63
    assert orc.conv_out_frames(200) == 85


def test_frame_counts():
    assert [int(orc.frame_count(s * 16000)) for s in (1, 4, 6, 30)] == [101, 401, 601, 3001]
    assert orc.N_FFT == 320 and orc.HOP == 160 and orc.N_BINS == 161


def test_window_matches_scipy():
    from scipy.signal import windows
    assert np.allclose(orc.hamming_window(320, True), windows.hamming(320, sym=True), atol=1e-15)
    assert np.allclose(orc.hamming_window(320, False), windows.hamming(320, sym=False), atol=1e-15)
    assert np.abs(orc.hamming_window(320, True) - orc.hamming_window(320, False)).max() > 5e-3


def test_mel_matches_torchaudio_slaney():
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(161, 0.0, 8000.0, 40, 16000, norm="slaney",
                                        mel_scale="slaney").T.double().numpy()
    mel = orc.mel_filterbank()
    assert mel.shape == (40, 161)
    assert np.abs(mel - ref).max() < 1e-7          # torchaudio builds it in float32
    nz = mel > 0
    assert nz.sum(axis=0).max() <= 2               # <= 2 filters per bin (banded-2)
    assert not nz[:, 0].any() and not nz[:, 160].any()


def test_stft_matches_torch_stft():
    b = _synth.make_batch(2, 3000, seed=3, ragged=True)
    win = orc.hamming_window()
    for i in range(2):
        li = int(b["lengths"][i])
        spec = orc.stft_frames(b["wave"][i], li, win)
        ref = torch.stft(torch.from_numpy(b["wave"][i, :li]).double(), 320, hop_length=160,
                         win_length=320, window=torch.from_numpy(win), center=True,
                         pad_mode="reflect", return_complex=True).numpy()
        assert spec.shape == ref.shape == (161, 1 + li // 160)
        assert np.abs(spec - ref).max() < 1e-10


def test_numpy_and_torch_twins_agree():
    b = _synth.make_batch(3, 2500, seed=21, ragged=True)
    for mm, cm in (("reim", "per_bin"), ("power", "global"), ("none", "none")):
        mr = b["mask_r"] if mm != "none" else None
        mi = b["mask_i"] if mm == "reim" else None
        z, fl = orc.lmfb_forward(b["wave"], b["lengths"], mr, mi, mask_mode=mm, cmvn_mode=cm)
        zt, flt = orc.lmfb_forward_torch(
            torch.from_numpy(b["wave"]), b["lengths"],
            None if mr is None else torch.from_numpy(mr), None if mi is None else torch.from_numpy(mi),
            mask_mode=mm, cmvn_mode=cm)
        assert np.array_equal(fl, flt.numpy())
        assert np.abs(z - zt.numpy()).max() < 1e-9
        for i in range(3):
            assert np.all(z[i, :, fl[i]:] == 0.0)


def test_cmvn_statistics():
    b = _synth.make_batch(2, 4000, seed=2, ragged=True)
    z, fl = orc.lmfb_forward(b["wave"], b["lengths"], b["mask_r"], b["mask_i"])
    for i in range(2):
        zi = z[i, :, :fl[i]]
        assert np.abs(zi.mean(axis=1)).max() < 1e-12
        assert np.abs(zi.std(axis=1, ddof=1) - 1.0).max() < 1e-12


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_oracle_fixtures_regenerate(name):
    g = _load(f"oracle_lmfb_{name}.npz")
    b = _synth.make_batch(int(g["n"]), int(g["max_len"]), seed=int(g["seed"]),
                          ragged=bool(g["ragged"]), tonal=bool(g["tonal"]))
    mm, cm = str(g["mask_mode"]), str(g["cmvn"])
    mr = b["mask_r"] if mm != "none" else None
    mi = b["mask_i"] if mm == "reim" else None
    z, fl = orc.lmfb_forward(b["wave"], b["lengths"], mr, mi, mask_mode=mm, cmvn_mode=cm)
    assert np.array_equal(fl, g["frame_lens"])
    assert np.abs(z - g["z"]).max() < 1e-11


def test_gradients_against_finite_differences():
    b = _synth.make_batch(1, 800, seed=4)
    g = orc.lmfb_grads(b["wave"], b["lengths"], b["mask_r"], b["mask_i"], b["grad_out"])
    rs = np.random.RandomState(0)
    for _ in range(4):
        f, t = rs.randint(1, 160), rs.randint(0, b["tmax"])
        h = 1e-6
        mp, mm = b["mask_r"].astype(np.float64), b["mask_r"].astype(np.float64)
        mp[0, f, t] += h
        mm[0, f, t] -= h
        zp, _ = orc.lmfb_forward(b["wave"], b["lengths"], mp, b["mask_i"])
        zm, _ = orc.lmfb_forward(b["wave"], b["lengths"], mm, b["mask_i"])
        fd = ((zp - zm) * b["grad_out"]).sum() / (2 * h)
        assert abs(fd - g["grad_mask_r"][0, f, t]) < 1e-5 * max(1.0, abs(fd))


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference not mounted")
def test_live_reference_collate_when_mounted():
    import sys
    sys.path.insert(0, REF)
    try:
        import loader_functions as lf
    finally:
        sys.path.remove(REF)
    batch, _ = _rebuild_collate_batch()
    ref = lf._collate_fn([(torch.from_numpy(f), t) for f, t in batch])
    mine = orc.collate(batch)
    for a, b in zip(mine, ref):
        assert np.array_equal(a, b.numpy())


def _l1_batch():
    g = _load("ref_l1loss.npz")
    rs = np.random.RandomState(int(g["seed"]))
    n, c, tmax = 4, 40, 37
    a = rs.randn(n, c, tmax).astype(np.float32)
    b = rs.randn(n, c, tmax).astype(np.float32)
    mask = np.zeros((n, 1, tmax), dtype=np.uint8)
    for i, l in enumerate(g["lens"]):
        mask[i, :, l:] = 1
    return g, a, b, mask


def test_l1loss_mask_matches_live_reference_fixture():
    g, a, b, mask = _l1_batch()
    loss, n_element = orc.l1loss_mask(a, b, mask)
    assert n_element == int(g["n_element"]) == 37 + 30 + 22 + 9      # frames, not elements
    assert abs(loss - float(g["loss"])) < 2e-6 * abs(float(g["loss"]))
    # the no-op masked_fill: padded frames contribute (fixing it changes the value)
    fixed, _ = orc.l1loss_mask(a, b, mask, fix_masking=True)
    assert fixed < loss
    with pytest.raises(RuntimeError):
        bad = mask.copy(); bad[0, 0, 0] = 1
        orc.l1loss_mask(a, b, bad)
