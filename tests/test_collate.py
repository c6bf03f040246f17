"""Host-side batch layout against the reference's collate (committed fixture + live reference
when mounted) and the N>1 utterance sharding over gloo.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import lmfb_oracle as orc
from aas_enhancement_b200 import (collate_wave, collate_wave_paired, ctc_sizes, frame_count,
                                  shard_utterances)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _wave_batch(seed=7):
    rs = np.random.RandomState(seed)
    lens = [37 * 160 + 5, 52 * 160, 52 * 160 + 159, 11 * 160 + 80, 29 * 160, 52 * 160 + 1, 7]
    batch, paired = [], []
    for i, li in enumerate(lens):
        w = rs.randn(li).astype(np.float32)
        c = rs.randn(li).astype(np.float32)
        txt = [int(v) for v in rs.randint(1, 29, size=3 + i)]
        batch.append((torch.from_numpy(w), txt))
        paired.append((torch.from_numpy(w), txt, torch.from_numpy(c)))
    return lens, batch, paired


def test_collate_wave_mirrors_reference_layout():
    lens, batch, _ = _wave_batch()
    inputs, targets, pct, tsz, mask, lengths = collate_wave(batch)
    # the reference collate applied to features of the same lengths gives the same bookkeeping
    feats = [(np.zeros((40, frame_count(l)), np.float32), t) for (w, t), l in zip(batch, lens)]
    r_inputs, r_targets, r_pct, r_tsz, r_mask = orc.collate(feats)
    assert inputs.dtype == torch.float32 and inputs.shape == (7, max(lens))
    assert targets.dtype == torch.int32 and np.array_equal(targets.numpy(), r_targets)
    assert pct.dtype == torch.float32 and np.array_equal(pct.numpy(), r_pct)
    assert tsz.dtype == torch.int32 and np.array_equal(tsz.numpy(), r_tsz)
    assert mask.dtype == torch.uint8 and np.array_equal(mask.numpy(), r_mask)
    assert lengths.dtype == torch.int32
    t = [frame_count(int(l)) for l in lengths]
    assert t == sorted(t, reverse=True)
    for i in range(7):
        li = int(lengths[i])
        assert torch.all(inputs[i, li:] == 0)
        assert mask[i, 0, :t[i]].sum() == 0 and mask[i, 0, t[i]:].all()


def test_collate_wave_paired_order_and_padding():
    lens, _, paired = _wave_batch()
    inputs, outputs, mask, targets, pct, tsz, lengths = collate_wave_paired(paired)
    i2, t2, p2, s2, m2, l2 = collate_wave([(a, b) for a, b, _ in paired])
    assert torch.equal(inputs, i2) and torch.equal(targets, t2) and torch.equal(pct, p2)
    assert torch.equal(tsz, s2) and torch.equal(mask, m2) and torch.equal(lengths, l2)
    assert outputs.shape == inputs.shape
    for i in range(len(lens)):
        assert torch.all(outputs[i, int(lengths[i]):] == 0)


def test_ctc_sizes_bit_exact_vs_reference_fixture():
    rows = np.load(os.path.join(GOLD, "ref_ctc_sizes.npz"))["rows"]
    for t, tmax, t_out, want in rows:
        pct = torch.FloatTensor(1)
        pct[0] = int(t) / float(int(tmax))
        assert int(ctc_sizes(pct, int(t_out))[0]) == want
    pct = torch.FloatTensor([135 / float(435)])
    assert int(ctc_sizes(pct, 203)[0]) == 62          # not floor(135*203/435) == 63


def test_shard_utterances_balances_frames():
    counts = [len(shard_utterances(30, 8, r)) for r in range(8)]
    assert counts == [4, 4, 4, 4, 4, 4, 3, 3]
    seen = sorted(i for r in range(8) for i in shard_utterances(30, 8, r))
    assert seen == list(range(30))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lens, batch, _ = _wave_batch(seed=3)
    inputs, targets, pct, tsz, mask, lengths = collate_wave(batch)
    mine = shard_utterances(inputs.shape[0], world, rank)
    frames = torch.tensor([sum(frame_count(int(lengths[i])) for i in mine)], dtype=torch.int64)
    owned = torch.zeros(inputs.shape[0], dtype=torch.int64)
    owned[mine] = 1
    dist.all_reduce(owned)                 # bookkeeping only: the data path has no collective
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, frames)
    ret[rank] = (owned.tolist(), [int(g) for g in gathered])
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    owned0, frames0 = ret[0]
    owned1, frames1 = ret[1]
    assert owned0 == owned1 == [1] * 7       # every utterance owned exactly once
    assert frames0 == frames1
    total = sum(frames0)
    assert max(frames0) - min(frames0) <= 53  # at most one utterance of imbalance
    assert total == sum(frame_count(l) for l in _wave_batch(seed=3)[0])


def test_ctc_boundary_matches_a_brute_force_sum():
    """CTCLoss (stand-in for warp-ctc at trainer_AAS.py:168): softmax inside, blank 0, cost summed
    over the batch, lengths from ctc_sizes.  Checked against an explicit sum over alignments."""
    import itertools
    from aas_enhancement_b200 import CTCLoss
    torch.manual_seed(2)
    t_len, n, c = 5, 2, 4
    acts = torch.randn(t_len, n, c, requires_grad=True)
    targets = [[1, 2], [3]]
    pct = torch.tensor([1.0, 0.8], dtype=torch.float32)
    sizes = ctc_sizes(pct, t_len)
    assert sizes.tolist() == [5, 4]
    loss = CTCLoss()(acts, torch.IntTensor([1, 2, 3]), sizes, torch.IntTensor([2, 1]))
    assert loss.shape == (1,)
    logp = torch.log_softmax(acts.detach().double(), dim=2)
    total = 0.0
    for b in range(n):
        tb = int(sizes[b])
        p = 0.0
        for path in itertools.product(range(c), repeat=tb):
            col = [k for k, g in itertools.groupby(path)]
            if [k for k in col if k != 0] == targets[b]:
                p += float(torch.exp(sum(logp[t, b, path[t]] for t in range(tb))))
        total += -np.log(p)
    assert abs(float(loss) - total) < 1e-4 * abs(total)
    loss.backward()
    assert torch.isfinite(acts.grad).all() and float(acts.grad[4, 1].abs().sum()) == 0.0   # beyond act_lens
