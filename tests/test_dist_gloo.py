"""Data-parallel plumbing on CPU: world_size-2 gloo processes (no GPU).  Covers the gradient bucket
all-reduce, parameter broadcast, the globally normalised advantages and the environment sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fullysnnforleggedrobots_b200.dist import GradientAllReducer, normalize_advantages_global, shard_range
    red = GradientAllReducer()
    torch.manual_seed(100 + rank)                 # different init per rank -> broadcast must equalise
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ELU(), torch.nn.Linear(7, 1))
    red.broadcast_parameters(model)
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(12, 5, generator=g)
    y_all = torch.randn(12, 1, generator=g)
    lo, hi = shard_range(12, rank, world)
    loss = ((model(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean()
    loss.backward()
    red.allreduce_gradients(model)
    adv_all = torch.randn(4, 12, generator=g)
    adv = normalize_advantages_global(adv_all[:, lo:hi].contiguous(), red.allreduce_sum)
    # the peer-memory gradient sum is a CUDA / NCCL feature: on a CPU (gloo) job every rank declines it alike and the
    # all-reduce stays a torch.distributed call
    p2p = red.enable_p2p(1000)
    torch.save({"params": [p.detach().clone() for p in model.parameters()],
                "grads": [p.grad.clone() for p in model.parameters()], "adv": adv, "range": (lo, hi),
                "p2p": (p2p, red.p2p is None, red.p2p_note)},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    # identical parameters and identical (averaged) gradients on both ranks
    for a, b in zip(r[0]["params"], r[1]["params"]):
        assert torch.equal(a, b)
    for a, b in zip(r[0]["grads"], r[1]["grads"]):
        assert torch.allclose(a, b, atol=0, rtol=0)
    # ... equal to the single-process gradient of the full-batch mean loss (equal shard sizes)
    model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ELU(), torch.nn.Linear(7, 1))
    with torch.no_grad():
        for p, q in zip(model.parameters(), r[0]["params"]):
            p.copy_(q)
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(12, 5, generator=g)
    y_all = torch.randn(12, 1, generator=g)
    ((model(x_all) - y_all) ** 2).mean().backward()
    for p, gr in zip(model.parameters(), r[0]["grads"]):
        assert torch.allclose(p.grad, gr, atol=1e-6, rtol=1e-5)
    # advantages normalised with GLOBAL statistics == single-process normalisation of the whole rollout
    adv_all = torch.randn(4, 12, generator=g)
    ref = (adv_all - adv_all.mean()) / (adv_all.std() + 1e-8)
    got = torch.cat([r[0]["adv"], r[1]["adv"]], dim=1)
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)
    assert r[0]["range"] == (0, 6) and r[1]["range"] == (6, 12)
    for i in range(world):
        ok, none, note = r[i]["p2p"]
        assert ok is False and none and "disabled" in note


def test_shard_range_covers_everything():
    from fullysnnforleggedrobots_b200.dist import shard_range
    for n in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
