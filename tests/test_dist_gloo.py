"""world_size-2 gloo tests (CPU) of the multi-GPU learner conventions: env-sharded ranks must reproduce
the single-process result for (a) advantage normalisation, (b) the PPO gradient after the flat-bucket all-reduce,
(c) the adaptive-KL learning rate."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _make_problem():
    from oracle.ppo_oracle import ActorCriticOracle
    torch.manual_seed(0)
    ac = ActorCriticOracle(num_obs=6, num_priv=2, num_hist=40, num_actions=3)
    M = 64
    g = torch.Generator().manual_seed(1)
    data = dict(hist=torch.randn(M, 40, generator=g), priv=torch.randn(M, 2, generator=g), actions=torch.randn(M, 3, generator=g),
                adv_raw=torch.randn(M, generator=g) * 2 + 0.3, returns=torch.randn(M, 1, generator=g), old_v=torch.randn(M, 1, generator=g),
                old_logp=torch.randn(M, generator=g) * 0.1 - 3.0, old_mu=torch.randn(M, 3, generator=g) * 0.1, old_sigma=torch.ones(M, 3))
    return ac, data


def _loss(ac, d, adv, scale):
    """sum over the local samples of the per-sample PPO loss terms, scaled by `scale` (= 1/global batch)."""
    dist_ = ac.dist(d["hist"])
    logp = dist_.log_prob(d["actions"]).sum(-1)
    ratio = torch.exp(logp - d["old_logp"])
    surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).sum() * scale
    v = ac.value(d["hist"], d["priv"])
    vc = d["old_v"] + (v - d["old_v"]).clamp(-0.2, 0.2)
    vl = torch.max((v - d["returns"]).pow(2), (vc - d["returns"]).pow(2)).sum() * scale
    ent = dist_.entropy().sum(-1).sum() * scale
    kl = torch.sum(torch.log(dist_.stddev / d["old_sigma"] + 1e-5) + (d["old_sigma"] ** 2 + (d["old_mu"] - dist_.mean) ** 2) / (2 * dist_.stddev ** 2) - 0.5, -1).sum() * scale
    return surr + vl - 0.01 * ent, kl.detach()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from go1_b200.dist_utils import global_advantage_stats, allreduce_flat_grads, adaptive_lr
    ac, d = _make_problem()
    M = d["hist"].shape[0]
    sl = slice(rank * M // world, (rank + 1) * M // world)
    local = {k: v[sl] for k, v in d.items()}
    mean, std = global_advantage_stats(local["adv_raw"])
    adv = (local["adv_raw"] - mean) / (std + 1e-8)
    loss, kl = _loss(ac, local, adv, 1.0 / M)
    loss.backward()
    flat = allreduce_flat_grads(list(ac.parameters()))
    dist.all_reduce(kl)
    lr = adaptive_lr(1e-3, float(kl))
    out[rank] = (mean, std, flat.numpy().copy(), float(kl), lr)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world = 2
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ac, d = _make_problem()
    a = d["adv_raw"]
    mean, std = float(a.mean()), float(a.std())
    adv = (a - a.mean()) / (a.std() + 1e-8)
    loss, kl = _loss(ac, d, adv, 1.0 / a.numel())
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in ac.parameters()]).numpy()
    from go1_b200.dist_utils import adaptive_lr
    for r in range(world):
        m, s, g, k, lr = out[r]
        assert abs(m - mean) < 1e-6 and abs(s - std) < 1e-6
        assert np.allclose(g, flat, rtol=1e-4, atol=1e-6)
        assert abs(k - float(kl)) < 1e-5 and lr == adaptive_lr(1e-3, float(kl))
    assert np.array_equal(out[0][2], out[1][2])          # identical bucket on every rank -> identical Adam step
