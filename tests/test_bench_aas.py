"""bench_aas.py (BASELINE.json configs[2]): the stand-in models have the reference's shapes, the flat
gradient buffer really is the models' .grad storage, and a two-rank gloo all-reduce of it averages the
ranks' gradients (the N > 1 path of the step, exercised on CPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_aas  # noqa: E402


def test_stand_in_models_have_the_reference_shapes():
    """Parameter counts of the reference's models at the README setting (counted by
    instantiating model.py's classes): stackedBRNN(40, 40, 500, 4) = 16,040,540; DeepSpeech(map 128,
    k 11, 5 x 1000) = 73,300,312."""
    d = bench_aas.StackedBRNN(40, 40)
    a = bench_aas.DeepSpeechAM()
    assert sum(p.numel() for p in d.parameters()) == 16_040_540
    assert sum(p.numel() for p in a.parameters()) == 73_300_312
    x = torch.randn(2, 40, 101)
    assert d(x).shape == (2, 40, 101)
    assert a(x).shape == (2, bench_aas.conv_out_frames(101), 29)
    assert bench_aas.conv_out_frames(200) == 85                 # T = 200 -> 85
    g = bench_aas.StackedBRNN(40, 2 * 161)
    assert g(x).shape == (2, 322, 101)                          # mask head: real rows, then imaginary rows


def test_flat_grads_are_the_parameters_grads():
    torch.manual_seed(0)
    d = bench_aas.StackedBRNN(40, 40, n_hidden=16, n_layers=2)
    a = bench_aas.DeepSpeechAM(n_map=8, n_hidden=12, n_layers=2)
    fg = bench_aas.FlatGrads([d, a], torch.device("cpu"))
    x = torch.randn(3, 40, 64)
    (d(x).sum() + a(x).sum()).backward()
    assert fg.flat.numel() == sum(p.numel() for m in (d, a) for p in m.parameters())
    o = 0
    for p in fg.params:
        assert p.grad.data_ptr() == fg.flat[o:].data_ptr()
        o += p.numel()
    assert float(fg.flat.abs().sum()) > 0
    fg.zero(d)
    a0, a1 = fg.spans[id(d)]
    assert float(fg.flat[a0:a1].abs().sum()) == 0 and float(fg.flat[a1:].abs().sum()) > 0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    d = bench_aas.StackedBRNN(40, 40, n_hidden=16, n_layers=1)
    fg = bench_aas.FlatGrads([d], torch.device("cpu"))
    torch.manual_seed(100 + rank)
    x = torch.randn(2, 40, 32)
    d(x).sum().backward()
    local = fg.flat.clone()
    dist.all_reduce(fg.flat, op=dist.ReduceOp.SUM)
    fg.flat /= world
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.allclose(fg.flat, sum(gathered) / world, atol=1e-6)
    if rank == 0:
        open(out, "w").write("ok" if ok else "mismatch")
    dist.destroy_process_group()


def test_two_rank_gloo_all_reduce_of_the_flat_buffer(tmp_path):
    out = str(tmp_path / "res.txt")
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
