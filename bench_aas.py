#!/usr/bin/env python3
"""BASELINE.json configs[2]: the full AAS training step around the LMFB front-end, utterance-sharded
data-parallel (one process per GPU, NCCL over NVLink for the gradient all-reduce only).

    python bench.py --workload aas_step_30x6s [--gpus N ...]          (bench.py dispatches here)

The step is trainer_AAS.py:131-194 with the front-end INSIDE it (SURVEY 3(A) slot #2, 8(e) row 2):

    x_ny      = LMFB(wave_ny)                          unmasked forward      (no grad)        [front-end]
    Mr, Mi    = G(x_ny)                                4 x 500 residual BLSTM + mask head     [torch / cuDNN]
    enhanced  = LMFB(wave_ny, Mr, Mi)                  masked forward + CMVN (autograd)       [front-end]
    G-step    : L1Loss_mask(D(enhanced), enhanced)  * w_adv          .backward(retain_graph)  (D grads dropped)
    D-step    : L1Loss_mask(D(enh.detach()), ...)   * (-kt) * w_adv  .backward()
    CTC       : CTCLoss(ASR(enhanced)) / N * w_ac                    .backward()   -> LMFB backward -> G
    clean     : x_cl = LMFB(wave_cl);  L1Loss_mask(D(x_cl), x_cl) * w_adv .backward()         [front-end]
    all-reduce of the G, D and ASR gradients (ONE flat fp32 buffer, NCCL), three Adam(amsgrad) steps,
    kt update from the two losses averaged over the ranks (2-float all-reduce).

G, D (stackedBRNN, model.py:203-252) and the acoustic model (DeepSpeech, model.py:256-335: 2 x Conv1d k11
+ BN + LeakyReLU, 5 x 1000 BatchRNN, BN + Linear) are dense cuDNN / cuBLAS work and OUT OF SCOPE of the
hand-written path; they are re-declared here from the reference's shapes as bench-side
stand-ins (random init; /root/reference does not exist on the GPU box) so that the front-end can be
measured inside the step it serves.  Under sharding the acoustic model's BatchNorm layers become
SyncBatchNorm (model.py:72, :290, :316).
"""
from __future__ import annotations

import json
import os
import statistics
import time
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

HOP, SR, N_BINS, N_MELS = 160, 16000, 161, 40
LABELS = "_'ABCDEFGHIJKLMNOPQRSTUVWXYZ "            # Common/labels.json: 29 symbols, blank first


# --------------------------------------------------------------------------- stand-in models
class _BiRNN(nn.Module):
    """model.py:66-105 (BRNN / BatchRNN): a bias-free bidirectional LSTM whose two directions are
    SUMMED; BatchRNN puts a sequence-wise BatchNorm1d in front."""

    def __init__(self, n_in, n_hidden, batch_norm=False):
        super().__init__()
        self.bn = nn.BatchNorm1d(n_in) if batch_norm else None
        self.rnn = nn.LSTM(n_in, n_hidden, bidirectional=True, bias=False)

    def forward(self, x):                                   # (T, N, H)
        if self.bn is not None:
            t, n = x.shape[0], x.shape[1]
            x = self.bn(x.reshape(t * n, -1)).view(t, n, -1)
        y, _ = self.rnn(x)
        return y.view(y.shape[0], y.shape[1], 2, -1).sum(2)


class StackedBRNN(nn.Module):
    """model.py:203-252: conv1x1 (I -> H), four residual bidirectional LSTMs, conv1x1 (H -> O)."""

    def __init__(self, n_in, n_out, n_hidden=500, n_layers=4):
        super().__init__()
        self.first_linear = nn.Conv1d(n_in, n_hidden, 1)
        self.rnns = nn.ModuleList([_BiRNN(n_hidden, n_hidden) for _ in range(n_layers)])
        self.final_linear = nn.Conv1d(n_hidden, n_out, 1)

    def forward(self, x):                                   # (N, I, T)
        h = self.first_linear(x).permute(2, 0, 1)           # (T, N, H)
        for r in self.rnns:
            h = r(h) + h
        return self.final_linear(h.permute(1, 2, 0))        # (N, O, T)


class DeepSpeechAM(nn.Module):
    """model.py:256-335 at the README setting (README.md:31, :72): map 128, kernel 11, stride 2 then 1,
    5 x 1000 BatchRNN (the first without BatchNorm), BatchNorm + bias-free Linear to the 29 labels.
    ``LeakyReLU(map)`` is the reference's literal call (model.py:291: the slope argument is `map`)."""

    def __init__(self, n_freq=N_MELS, n_map=128, kernel=11, stride=2, n_hidden=1000, n_layers=5, n_classes=len(LABELS)):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv1d(n_freq, n_map, kernel, stride=stride), nn.BatchNorm1d(n_map), nn.LeakyReLU(n_map, inplace=True),
            nn.Conv1d(n_map, n_map, kernel, stride=1), nn.BatchNorm1d(n_map), nn.LeakyReLU(n_map, inplace=True))
        rnns = [("0", _BiRNN(n_map, n_hidden, batch_norm=False))]
        rnns += [(str(i), _BiRNN(n_hidden, n_hidden, batch_norm=True)) for i in range(1, n_layers)]
        self.rnns = nn.Sequential(OrderedDict(rnns))
        self.fc_bn = nn.BatchNorm1d(n_hidden)
        self.fc = nn.Linear(n_hidden, n_classes, bias=False)

    def forward(self, x):                                   # (N, 40, T) -> (N, T', C)
        h = self.conv(x).permute(2, 0, 1)                   # (T', N, map)
        h = self.rnns(h)
        t, n = h.shape[0], h.shape[1]
        y = self.fc(self.fc_bn(h.reshape(t * n, -1))).view(t, n, -1)
        return y.transpose(0, 1)


def conv_out_frames(t, kernel=11, stride=2):                # model.py:289-297 geometry
    return ((t - kernel) // stride + 1) - kernel + 1


class FlatGrads:
    """All gradients of several modules as views into ONE flat buffer: a single NCCL all-reduce per step
    (the reference accumulates over four backward calls per step, trainer_AAS.py:150-180, so the
    reduction must happen once, after the last of them)."""

    def __init__(self, modules, device):
        self.params = [p for m in modules for p in m.parameters() if p.requires_grad]
        self.spans = {}
        total = 0
        for m in modules:
            n = sum(p.numel() for p in m.parameters() if p.requires_grad)
            self.spans[id(m)] = (total, total + n)
            total += n
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self, module=None):
        if module is None:
            self.flat.zero_()
        else:
            a, b = self.spans[id(module)]
            self.flat[a:b].zero_()


class AASStep:
    def __init__(self, dev, n, samples, world, seed):
        from aas_enhancement_b200 import LMFBFrontEnd, L1Loss_mask, CTCLoss, ctc_sizes
        self.dev, self.n, self.samples, self.world = dev, n, samples, world
        self.tmax = 1 + samples // HOP
        torch.manual_seed(seed)
        self.fe_plain = LMFBFrontEnd(mask_mode="none", cmvn_mode="per_bin").to(dev)
        self.fe_mask = LMFBFrontEnd(mask_mode="reim", cmvn_mode="per_bin").to(dev)
        self.G = StackedBRNN(N_MELS, 2 * N_BINS).to(dev)             # mask head: 161 real + 161 imaginary rows (model.py:164-165)
        self.D = StackedBRNN(N_MELS, N_MELS).to(dev)
        self.ASR = DeepSpeechAM().to(dev)
        with torch.no_grad():                                        # masks around 1: the enhanced features stay well-conditioned
            self.G.final_linear.bias.fill_(1.0)
            self.G.final_linear.weight.mul_(0.1)
        if world > 1:
            self.ASR = nn.SyncBatchNorm.convert_sync_batchnorm(self.ASR)
        self.grads = FlatGrads([self.G, self.D, self.ASR], dev)
        kw = dict(lr=1e-4, betas=(0.5, 0.999), amsgrad=True)          # trainer_AAS.py:127-129
        self.opts = [torch.optim.Adam(m.parameters(), **kw) for m in (self.G, self.D, self.ASR)]
        self.l1, self.ctc, self.ctc_sizes = L1Loss_mask(), CTCLoss(), ctc_sizes
        self.kt, self.lambda_k, self.gamma = 0.0, 0.001, 0.5          # config.py BEGAN defaults
        self.w_adv, self.w_ac = 1.0, 1.0
        gen = torch.Generator(device=dev)
        gen.manual_seed(seed)
        self.wave_ny = (0.1 * torch.randn(n, samples, generator=gen, device=dev)).clamp_(-1, 1)
        self.wave_cl = (0.1 * torch.randn(n, samples, generator=gen, device=dev)).clamp_(-1, 1)
        self.lengths = torch.full((n,), samples, dtype=torch.int32, device=dev)
        self.mask = torch.zeros(n, 1, self.tmax, dtype=torch.uint8, device=dev)      # 1 = padding (loader_functions.py:60)
        self.pct = torch.ones(n, dtype=torch.float32)                                 # input_percentages (:57)
        rs = np.random.RandomState(seed)
        lens = rs.randint(20, 60, size=n)
        self.target_sizes = torch.from_numpy(lens.astype(np.int32))
        self.targets = torch.from_numpy(rs.randint(1, len(LABELS), size=int(lens.sum())).astype(np.int32))
        self.n_params = {k: sum(p.numel() for p in m.parameters()) for k, m in (("G", self.G), ("D", self.D), ("ASR", self.ASR))}
        self.ev = None

    # one training step; `sync` = all-reduce the gradients (False: timing the step without communication)
    def step(self, sync=True, wave_ny=None, wave_cl=None):
        import torch.distributed as dist
        wave_ny = self.wave_ny if wave_ny is None else wave_ny
        wave_cl = self.wave_cl if wave_cl is None else wave_cl
        n = self.n
        self.grads.zero()
        with torch.no_grad():
            x_ny, _ = self.fe_plain(wave_ny, self.lengths)
        m = self.G(x_ny)
        enhanced, _ = self.fe_mask(wave_ny, self.lengths, m[:, :N_BINS], m[:, N_BINS:])
        enhanced_d = enhanced.detach()
        # adversarial: G-step (trainer_AAS.py:145-153)
        l_g, _ = self.l1(self.D(enhanced), enhanced, self.mask)
        l_g = l_g * self.w_adv
        l_g.backward(retain_graph=True)
        self.grads.zero(self.D)                                      # "this makes no gradient for discriminator"
        # adversarial: D-step (:155-161)
        l_d, _ = self.l1(self.D(enhanced_d), enhanced_d, self.mask)
        (l_d * (-self.kt) * self.w_adv).backward()
        # CTC (:163-172)
        prob = self.ASR(enhanced).transpose(0, 1)                    # (T', N, C)
        sizes = self.ctc_sizes(self.pct, prob.shape[0])
        l_ctc = self.w_ac * self.ctc(prob, self.targets, sizes, self.target_sizes) / n
        l_ctc.backward()
        # clean stream (:174-182)
        with torch.no_grad():
            x_cl, _ = self.fe_plain(wave_cl, self.lengths)
        l_c, _ = self.l1(self.D(x_cl), x_cl, self.mask)
        l_c = l_c * self.w_adv
        l_c.backward()
        # one all-reduce for all three models (421 MB fp32), then the three optimiser steps (:185-188)
        if self.ev is not None:
            self.ev[0].record()
        if sync and self.world > 1:
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.AVG)
        if self.ev is not None:
            self.ev[1].record()
        for o in self.opts:
            o.step()
        # proportional control (:190-194): both losses averaged over the ranks so that kt stays identical
        pair = torch.stack([l_c.detach().reshape(()), l_g.detach().reshape(())])
        if sync and self.world > 1:
            dist.all_reduce(pair, op=dist.ReduceOp.AVG)
        return pair, l_ctc.detach()

    def update_kt(self, pair):
        l_cl, l_ny = (float(v) for v in pair.tolist())               # the step's device -> host read
        self.kt = max(min(1.0, self.kt + self.lambda_k * (self.gamma * l_cl - l_ny)), 0.0)

    # the front-end work of one step, in isolation (what the reference does on the host, offline)
    def frontend_only(self, masks, gout):
        mr, mi = masks
        with torch.no_grad():
            self.fe_plain(self.wave_ny, self.lengths)
            self.fe_plain(self.wave_cl, self.lengths)
        z, _ = self.fe_mask(self.wave_ny, self.lengths, mr, mi)
        z.backward(gout)
        mr.grad = None
        mi.grad = None


def _timed(fn, steps, dev, world):
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps


def run(args, world, rank, local, dev):
    import torch.distributed as dist
    import bench
    n, secs, _ = bench.WORKLOADS[args.workload]
    samples = int(secs * SR)
    steps = min(args.steps, 50)
    warm = max(min(args.warmup, 10), 3)
    job = AASStep(dev, n, samples, world, 123 + rank)

    def full():
        pair, _ = job.step(True)
        job.update_kt(pair)

    def nosync():
        pair, _ = job.step(False)
        job.update_kt(pair)

    for _ in range(warm):
        full()
    sampler = bench.ClockSampler(local)
    sampler.start()
    sampler.region(True)
    ms_step = _timed(full, steps, dev, world)
    sampler.region(False)
    clocks = sampler.stop()
    ms_nosync = _timed(nosync, max(steps // 2, 3), dev, world)
    # the all-reduce by itself, timed with events around the call inside the step
    job.ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    ar = []
    for _ in range(5):
        full()
        torch.cuda.synchronize()
        ar.append(job.ev[0].elapsed_time(job.ev[1]))
    job.ev = None
    ms_allreduce = statistics.median(ar)
    if world > 1:
        t = torch.tensor([ms_allreduce], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_allreduce = float(t.item())
    # ... and by itself, all ranks released together by a barrier (the in-step figure also contains the wait for
    # the slowest rank of a 590 ms step)
    ms_ar_iso = None
    if world > 1:
        iso = []
        for _ in range(5):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(job.grads.flat, op=dist.ReduceOp.AVG)
            e1.record()
            torch.cuda.synchronize()
            iso.append(e0.elapsed_time(e1))
        t = torch.tensor([statistics.median(iso)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_ar_iso = float(t.item())
    # the front-end work of the step in isolation
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    masks = (torch.rand(n, N_BINS, job.tmax, generator=gen, device=dev).requires_grad_(True),
             torch.rand(n, N_BINS, job.tmax, generator=gen, device=dev).requires_grad_(True))
    gout = torch.randn(n, N_MELS, job.tmax, generator=gen, device=dev)
    for _ in range(3):
        job.frontend_only(masks, gout)
    ms_fe = _timed(lambda: job.frontend_only(masks, gout), 20, dev, world)

    # e2e: both waves and the transcripts come from pinned host memory every step, the losses go back
    host = dict(ny=job.wave_ny.cpu().pin_memory(), cl=job.wave_cl.cpu().pin_memory())
    d_ny, d_cl = torch.empty_like(job.wave_ny), torch.empty_like(job.wave_cl)
    tg_host = job.targets.pin_memory()
    h2d = (host["ny"].numel() + host["cl"].numel()) * 4 + tg_host.numel() * 4 + job.target_sizes.numel() * 4

    def e2e_step():
        d_ny.copy_(host["ny"], non_blocking=True)
        d_cl.copy_(host["cl"], non_blocking=True)
        tg_host.to(dev, non_blocking=True)                           # (the CTC call itself takes host labels)
        pair, l_ctc = job.step(True, d_ny, d_cl)
        job.update_kt(pair)
        float(l_ctc)

    for _ in range(2):
        e2e_step()
    ms_e2e = _timed(e2e_step, max(steps // 2, 5), dev, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    audio_s = n * secs
    grad_bytes = job.grads.flat.numel() * 4
    line = {
        "metric": "LMFB fwd+bwd audio-seconds per second",
        "value": world * audio_s / (ms_step / 1e3), "unit": "audio-s/s", "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench.workload_config(args.workload, world),
        "clocks": clocks,
        "e2e": {"value": world * audio_s / (ms_e2e / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 12, "ms_per_step": ms_e2e,
                "api": "AAS step through LMFBFrontEnd / L1Loss_mask / CTCLoss; noisy + clean waves and "
                       "transcripts copied from pinned host memory and the three losses read back, every step"},
        "gpu_launches": 9 * steps,
        "gpu_launches_note": "front-end kernels only: 3 forward calls (K1 + CMVN), 1 backward (CMVN + K1), "
                             "plus L1Loss_mask kernels; the models are cuDNN / cuBLAS library calls",
        "aas_step": {
            "ms_step": ms_step, "ms_step_without_allreduce": ms_nosync, "ms_allreduce_in_step": ms_allreduce,
            "ms_frontend_isolated": ms_fe, "ms_models_and_losses": ms_nosync - ms_fe,
            "frontend_share_of_step": ms_fe / ms_step,
            "allreduce_bytes": grad_bytes,
            "ms_allreduce_isolated": ms_ar_iso,
            "allreduce_busbw_gbs": (2 * (world - 1) / world) * grad_bytes / (ms_ar_iso / 1e3) / 1e9 if ms_ar_iso else None,
            "params": job.n_params, "nccl_ranks": world,
            "sync_batchnorm": world > 1,
            "models": "stand-ins re-declared from model.py:203-252 (G with a 322-row mask head, D) and :256-335 "
                      "(acoustic model), random init; cuDNN LSTMs in fp32 (TF32 off)"},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, world):
    """The reference's own CPU implementation of the PATH inside this step: the three front-end passes
    (noisy plain forward, masked forward + backward, clean forward) on the host cores.  The models are
    not part of the path (in the reference they run on the GPU too)."""
    import bench
    n, secs, _ = bench.WORKLOADS[args.workload]
    samples = int(secs * SR)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_s = bench.cpu_sample_size(n)
    from oracle import lmfb_torch_cpu as cpu
    wave, wave_c, mr, mi, g, mel, win = bench._cpu_inputs(n_s, samples)

    def step():
        with torch.no_grad():
            cpu.forward_batched(wave, None, None, mel, win, "per_bin", "none")
            cpu.forward_batched(wave_c, None, None, mel, win, "per_bin", "none")
        cpu.fwd_bwd_batched(wave, mr, mi, g, mel, win, "per_bin", "reim")

    for _ in range(max(args.warmup, 3)):
        step()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    value = n_s * secs / med
    sample = (f"front-end passes of the step only (noisy plain forward, masked forward + backward, clean forward), "
              f"{n_s} of {n} utterances x {secs:g} s, torch CPU path, {cores} threads, median of {len(times)}")
    print(json.dumps({
        "impl": "reference", "metric": "LMFB fwd+bwd audio-seconds per second", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": med * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench.workload_config(args.workload, world),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))
