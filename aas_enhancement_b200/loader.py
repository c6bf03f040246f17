"""Wave-level dataset / sampler / loader: the host-side mirror of the reference's data plumbing
with waveforms in place of precomputed features (8(f) rank 3).

Reference                                             here
--------------------------------------------------   ------------------------------------------
FeatDataset        (loader_functions.py:11-44)        WaveDataset   same manifest format
                                                      ``path,transcript[,paired path]``, same
                                                      ``parse_transcript`` (incl. its quirk: label
                                                      index 0 is filtered out), items are 1-D waves
convert_numpy_to_pytorch.py:20-26 (.npy -> .pt7)      save_wave / load_wave: 1-D float32 (or int16)
                                                      tensors written with ``torch.save``, also
                                                      ``.npy``; int16 is scaled by 1/32768 on load
FeatSampler        (loader_functions.py:118-137)      WaveSampler   identical bucketing: consecutive
                                                      ids in bins of ``batch_size``, ids shuffled
                                                      inside a bin, ``shuffle()`` permutes the bins
FeatLoader(_paired)(loader_functions.py:107-116)      WaveLoader(_paired): torch DataLoader whose
                                                      collate_fn is collate_wave(_paired); pinned
                                                      in the main process (pin_memory=True)
DataLoader.next    (data_loader.py:8-83)              WaveDataLoader.next(cl_ny, type): same five
                                                      streams, same re-shuffle / restart on exhaustion
_get_variable_*    (utils.py:139-160)                 to_device: one asynchronous H2D copy per tensor
                                                      from the pinned batch on a caller-given stream

Everything here is host code (the reference's is Python too); the arithmetic is in the kernels.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.sampler import Sampler

from .collate import collate_wave, collate_wave_paired


def save_wave(path: str, wave, int16: bool = False) -> None:
    """Write a mono waveform as a 1-D tensor with ``torch.save`` (the reference stores its
    features the same way, data/convert_numpy_to_pytorch.py:20-26).  ``int16=True`` halves the file
    and the bytes on the wire; samples are expected in [-1, 1)."""
    w = torch.as_tensor(np.asarray(wave), dtype=torch.float32).reshape(-1)
    if int16:
        w = torch.clamp(torch.round(w * 32768.0), -32768, 32767).to(torch.int16)
    torch.save(w, path)


def load_wave(path: str) -> torch.Tensor:
    """1-D float32 waveform from a ``torch.save``'d tensor or a ``.npy`` file (int16 -> /32768)."""
    if path.endswith(".npy"):
        w = torch.from_numpy(np.load(path))
    else:
        w = torch.load(path)
    w = w.reshape(-1)
    if w.dtype == torch.int16:
        return w.to(torch.float32) / 32768.0
    return w.to(torch.float32)


class WaveDataset(Dataset):
    """``manifest``: one ``wave_path,transcript_path[,paired_wave_path]`` per line
    (loader_functions.py:12-16).  Items are ``(wave, transcript)`` or
    ``(wave, transcript, paired_wave)`` exactly like FeatDataset.__getitem__ (:21-35)."""

    def __init__(self, manifest, labels):
        with open(manifest) as f:
            ids = f.readlines()
        self.ids = [x.strip().split(',') for x in ids if x.strip()]
        self.size = len(self.ids)
        self.labels_map = dict([(labels[i], i) for i in range(len(labels))])
        super().__init__()

    def __getitem__(self, index):
        sample = self.ids[index]
        wave = load_wave(sample[0])
        transcript = self.parse_transcript(sample[1])
        if len(sample) == 2:
            return wave, transcript
        return wave, transcript, load_wave(sample[2])

    def parse_transcript(self, transcript_path):
        # loader_functions.py:37-41, literally: `filter(None, ...)` also drops label index 0
        with open(transcript_path, 'r', encoding='utf8') as transcript_file:
            transcript = transcript_file.read().replace('\n', '')
        return list(filter(None, [self.labels_map.get(x) for x in list(transcript)]))

    def __len__(self):
        return self.size


class WaveSampler(Sampler):
    """Batches of consecutive ids (the manifest is assumed sorted by size), shuffled inside a batch;
    ``shuffle()`` permutes the batches (loader_functions.py:118-137).  Uses ``numpy.random`` like
    the reference, so ``np.random.seed`` reproduces its order."""

    def __init__(self, data_source, batch_size=1):
        self.data_source = data_source
        ids = list(range(0, len(data_source)))
        self.bins = [ids[i:i + batch_size] for i in range(0, len(ids), batch_size)]

    def __iter__(self):
        for ids in self.bins:
            np.random.shuffle(ids)
            yield ids

    def __len__(self):
        return len(self.bins)

    def shuffle(self):
        np.random.shuffle(self.bins)


class WaveLoader(DataLoader):
    """torch DataLoader with ``collate_wave`` (module-level, so it pickles under spawn/forkserver).
    Nothing is pinned inside the collate function: with ``num_workers > 0`` it runs in a worker
    process, where a pinned allocation would initialise CUDA (or fail after a fork).  Pass
    ``pin_memory=True`` to let the DataLoader's pin thread do it in the main process, or let
    :func:`to_device` pin (it does when the batch is not pinned yet)."""

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("collate_fn", collate_wave)
        super().__init__(*args, **kwargs)


class WaveLoader_paired(DataLoader):
    def __init__(self, *args, **kwargs):
        kwargs.setdefault("collate_fn", collate_wave_paired)
        super().__init__(*args, **kwargs)


class WaveDataLoader:
    """data_loader.py:8-83 with waves: ``next(cl_ny, type)`` returns the next batch tuple of the
    requested stream ('ny'/'train', 'ny'/'trsub', 'ny'/'val', 'ny'/'val2', 'cl'/'train'); an
    exhausted training stream shuffles its sampler's bins and restarts, an exhausted evaluation
    stream restarts (the reference's stray ``loader = self.te_dl`` at :74 is not reproduced)."""

    def __init__(self, batch_size, paired=False, tr_cl_manifest="", tr_ny_manifest="", trsub_manifest="",
                 val_manifest="", val2_manifest="", labels=None, num_workers=0, pin_memory=None):
        self.batch_size = batch_size
        self.labels = labels
        self.num_workers = num_workers
        # pinned by the DataLoader's pin thread in the MAIN process (never inside a worker)
        self.pin_memory = torch.cuda.is_available() if pin_memory is None else bool(pin_memory)
        self.Loader = WaveLoader_paired if paired else WaveLoader
        self._ds, self._sp, self._it = {}, {}, {}
        for key, manifest, sampled in (("cl/train", tr_cl_manifest, True), ("ny/train", tr_ny_manifest, True),
                                       ("ny/trsub", trsub_manifest, False), ("ny/val", val_manifest, False),
                                       ("ny/val2", val2_manifest, False)):
            if len(manifest) > 0:
                self._ds[key] = WaveDataset(manifest=manifest, labels=labels)
                if sampled:
                    self._sp[key] = WaveSampler(self._ds[key], batch_size=batch_size)
                self._it[key] = self._make(key)

    def _make(self, key):
        if key in self._sp:
            return iter(self.Loader(self._ds[key], num_workers=self.num_workers, batch_sampler=self._sp[key],
                                    pin_memory=self.pin_memory))
        return iter(self.Loader(self._ds[key], batch_size=self.batch_size, num_workers=self.num_workers,
                                pin_memory=self.pin_memory))

    def next(self, cl_ny='', type=''):
        key = f"{cl_ny}/{type}"
        if key not in self._it:
            raise KeyError(f"no manifest was given for stream {key!r}")
        try:
            return next(self._it[key])
        except StopIteration:
            if key in self._sp:
                self._sp[key].shuffle()
            self._it[key] = self._make(key)
            return next(self._it[key])


def to_device(batch, device=None, stream=None):
    """utils.py:139-160 for a whole batch tuple: every tensor is copied host -> device
    asynchronously on ``stream`` (default: the current stream); non-tensors pass through.  Host
    tensors that are not pinned yet (``pin_memory=False`` loaders) are pinned here, in the calling
    process, so that every copy is a true asynchronous DMA."""
    device = torch.device("cuda") if device is None else torch.device(device)
    ctx = torch.cuda.stream(stream) if stream is not None else _Null()

    def move(t):
        if not isinstance(t, torch.Tensor):
            return t
        if device.type == "cuda" and not t.is_cuda and not t.is_pinned() and t.numel() > 0:
            t = t.pin_memory()
        return t.to(device, non_blocking=True)
    with ctx:
        return tuple(move(t) for t in batch)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
