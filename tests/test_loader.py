"""Wave-level dataset / sampler / loader (8(f) rank 3) against the reference's
FeatDataset / FeatSampler semantics (imported live when /root/reference is mounted).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from aas_enhancement_b200 import (WaveDataset, WaveSampler, WaveDataLoader, collate_wave, frame_count,
                                  load_wave, save_wave)

REF = "/root/reference/Speech_enhancement_by_AAS"
LABELS = "_'ABCDEFGHIJKLMNOPQRSTUVWXYZ "


def _ref_module():
    if not os.path.isdir(REF):
        return None
    sys.path.insert(0, REF)
    try:
        import loader_functions
        return loader_functions
    except Exception:
        return None
    finally:
        sys.path.remove(REF)


@pytest.fixture()
def corpus(tmp_path):
    rs = np.random.RandomState(3)
    lines, lines_paired, waves = [], [], []
    texts = ["HELLO WORLD", "A_B", "IT'S", "Z", "SPEECH ENHANCEMENT", "B200", "OK"]
    for i, txt in enumerate(texts):
        li = 160 * (5 + 3 * i) + 7 * i
        w = (0.3 * rs.randn(li)).clip(-1, 1).astype(np.float32)
        c = (0.3 * rs.randn(li + 11)).clip(-1, 1).astype(np.float32)
        wp, cp, tp = tmp_path / f"u{i}.pt", tmp_path / f"c{i}.npy", tmp_path / f"u{i}.txt"
        save_wave(str(wp), w, int16=(i % 2 == 1))
        np.save(str(cp), c)
        tp.write_text(txt + "\n", encoding="utf8")
        lines.append(f"{wp},{tp}")
        lines_paired.append(f"{wp},{tp},{cp}")
        waves.append(w)
    m, mp = tmp_path / "m.csv", tmp_path / "mp.csv"
    m.write_text("\n".join(lines) + "\n")
    mp.write_text("\n".join(lines_paired) + "\n")
    return dict(manifest=str(m), paired=str(mp), waves=waves, texts=texts)


def test_wave_files_round_trip(tmp_path):
    w = np.linspace(-1, 0.999, 1000).astype(np.float32)
    save_wave(str(tmp_path / "a.pt"), w)
    assert torch.equal(load_wave(str(tmp_path / "a.pt")), torch.from_numpy(w))
    save_wave(str(tmp_path / "b.pt"), w, int16=True)
    got = load_wave(str(tmp_path / "b.pt"))
    assert got.dtype == torch.float32 and float((got - torch.from_numpy(w)).abs().max()) <= 0.5 / 32768 + 1e-7
    np.save(str(tmp_path / "c.npy"), w.astype(np.float64))
    assert load_wave(str(tmp_path / "c.npy")).dtype == torch.float32


def test_dataset_items_and_transcripts_match_reference(corpus):
    ds = WaveDataset(corpus["manifest"], LABELS)
    assert len(ds) == 7
    ref = _ref_module()
    for i in range(7):
        wave, txt = ds[i]
        assert wave.dtype == torch.float32 and wave.dim() == 1 and wave.numel() == len(corpus["waves"][i])
        tol = 0.5 / 32768 + 1e-7 if i % 2 == 1 else 0.0
        assert float((wave - torch.from_numpy(corpus["waves"][i])).abs().max()) <= tol
        # the reference's parse: characters not in the label set AND label index 0 are dropped
        want = [LABELS.index(ch) for ch in corpus["texts"][i] if ch in LABELS and LABELS.index(ch) != 0]
        assert txt == want
    if ref is not None:                                   # live reference: same manifest, same parse
        rds = ref.FeatDataset.__new__(ref.FeatDataset)
        rds.labels_map = dict([(LABELS[i], i) for i in range(len(LABELS))])
        for i in range(7):
            assert ds[i][1] == rds.parse_transcript(ds.ids[i][1])
    dsp = WaveDataset(corpus["paired"], LABELS)
    w, t, c = dsp[2]
    assert c.numel() == w.numel() + 11 and t == ds[2][1]


def test_sampler_buckets_like_the_reference(corpus):
    ds = WaveDataset(corpus["manifest"], LABELS)
    np.random.seed(11)
    sp = WaveSampler(ds, batch_size=3)
    assert len(sp) == 3 and sorted(sum(sp.bins, [])) == list(range(7))
    assert sp.bins == [[0, 1, 2], [3, 4, 5], [6]]                            # loader_functions.py:126-127
    first = [list(b) for b in sp]
    assert [sorted(b) for b in first] == [[0, 1, 2], [3, 4, 5], [6]]          # consecutive ids per bin
    sp.shuffle()
    ref = _ref_module()
    if ref is not None:
        # the reference's constructor no longer runs (torch's Sampler.__init__ takes no argument
        # any more); its __iter__ / shuffle are run live on the same initial bins
        rsp = ref.FeatSampler.__new__(ref.FeatSampler)
        rsp.bins = [[0, 1, 2], [3, 4, 5], [6]]
        np.random.seed(11)
        rfirst = [list(b) for b in rsp]
        rsp.shuffle()
        assert first == rfirst and sp.bins == rsp.bins


def test_data_loader_streams_restart_and_layout(corpus):
    np.random.seed(5)
    dl = WaveDataLoader(batch_size=3, tr_ny_manifest=corpus["manifest"], tr_cl_manifest=corpus["manifest"],
                        val_manifest=corpus["manifest"], labels=LABELS)
    seen = 0
    for _ in range(3):                                    # one epoch of the noisy training stream
        inputs, targets, pct, tsz, mask, lengths = dl.next(cl_ny='ny', type='train')
        n = inputs.shape[0]
        seen += n
        t = [frame_count(int(l)) for l in lengths]
        assert t == sorted(t, reverse=True) and mask.shape == (n, 1, t[0]) and mask.dtype == torch.uint8
        assert pct.dtype == torch.float32 and tsz.dtype == torch.int32 and targets.dtype == torch.int32
        assert int(tsz.sum()) == targets.numel()
    assert seen == 7
    again = dl.next(cl_ny='ny', type='train')             # exhausted: bins reshuffled, stream restarts
    assert again[0].shape[0] in (1, 3)
    v1 = dl.next(cl_ny='ny', type='val'); dl.next(cl_ny='ny', type='val'); dl.next(cl_ny='ny', type='val')
    v4 = dl.next(cl_ny='ny', type='val')                  # evaluation stream restarts in file order
    assert torch.equal(v1[0], v4[0])
    assert dl.next(cl_ny='cl', type='train')[0].dim() == 2
    with pytest.raises(KeyError):
        dl.next(cl_ny='ny', type='val2')
    # the same batch through the collate function directly
    ds = WaveDataset(corpus["manifest"], LABELS)
    direct = collate_wave([ds[0], ds[1], ds[2]])
    assert torch.equal(v1[0], direct[0]) and torch.equal(v1[5], direct[5])


def test_paired_loader_tuple_order(corpus):
    dl = WaveDataLoader(batch_size=4, paired=True, trsub_manifest=corpus["paired"], labels=LABELS)
    inputs, outputs, mask, targets, pct, tsz, lengths = dl.next(cl_ny='ny', type='trsub')
    assert inputs.shape == outputs.shape and mask.shape[0] == 4 and lengths.dtype == torch.int32
    for i in range(4):
        assert torch.all(outputs[i, int(lengths[i]):] == 0)


def test_loader_with_a_worker_process_and_main_process_pinning(corpus):
    """num_workers=1 is the reference's setting (data_loader.py:27).  The collate function runs in the
    worker and must not touch CUDA / pinned memory; the loaders pickle (spawn / forkserver)."""
    import pickle
    from aas_enhancement_b200 import WaveLoader, WaveLoader_paired, collate_wave_paired
    ds = WaveDataset(corpus["manifest"], LABELS)
    ld = WaveLoader(ds, batch_size=3, num_workers=1)
    pickle.dumps(ld.collate_fn)
    batches = list(ld)
    assert len(batches) == 3
    inputs, targets, pct, sizes, mask, lengths = batches[0]
    ref = collate_wave([ds[i] for i in range(3)])
    assert torch.equal(inputs, ref[0]) and torch.equal(pct, ref[2]) and torch.equal(mask, ref[4])
    assert not inputs.is_pinned()                      # pinning is the main process's job
    dsp = WaveDataset(corpus["paired"], LABELS)
    ldp = WaveLoader_paired(dsp, batch_size=2, num_workers=1)
    assert ldp.collate_fn is collate_wave_paired
    first = next(iter(ldp))
    assert len(first) == 7 and first[0].shape == first[1].shape
