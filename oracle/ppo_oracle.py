"""ppo_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT.

torch-CPU (autograd) restatement of go1_gym_learn/ppo_cse for the parity tests and for bench.py's CPU
baseline / `--impl reference` arm:
  ActorCritic      actor_critic.py:19-147   (3 ELU MLPs + state-independent std)
  GAE + normalise  rollout_storage.py:74-88
  update           ppo.py:97-205            (clipped surrogate + clipped value loss - entropy, adaptive-KL LR,
                                             clip_grad_norm_ 1.0, Adam; then the adaptation-module MSE step)
Pinned against tests/golden/ppo.npz, which was produced by the reference's own code (tests/test_ppo_oracle.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

ARGS = dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.01, num_learning_epochs=5,
            num_mini_batches=4, learning_rate=1e-3, adaptation_module_learning_rate=1e-3, gamma=0.99, lam=0.95,
            desired_kl=0.01, max_grad_norm=1.0)


def mlp(i, hidden, o):
    layers, d = [], i
    for h in hidden:
        layers += [nn.Linear(d, h), nn.ELU()]
        d = h
    return nn.Sequential(*layers, nn.Linear(d, o))


class ActorCriticOracle(nn.Module):
    def __init__(self, num_obs=70, num_priv=2, num_hist=2100, num_actions=12):
        super().__init__()
        self.adaptation_module = mlp(num_hist, [256, 128], num_priv)
        self.actor_body = mlp(num_hist + num_priv, [512, 256, 128], num_actions)
        self.critic_body = mlp(num_hist + num_priv, [512, 256, 128], 1)
        self.std = nn.Parameter(torch.ones(num_actions))

    def dist(self, hist):
        latent = self.adaptation_module(hist)
        mean = self.actor_body(torch.cat((hist, latent), -1))
        return torch.distributions.Normal(mean, mean * 0. + self.std)

    def value(self, hist, priv):
        return self.critic_body(torch.cat((hist, priv), -1))


def gae(rewards, dones, values, last_values, gamma=0.99, lam=0.95):
    T = rewards.shape[0]
    returns, adv = torch.zeros_like(rewards), 0
    for t in reversed(range(T)):
        nv = last_values if t == T - 1 else values[t + 1]
        nt = 1.0 - dones[t].float()
        delta = rewards[t] + nt * gamma * nv - values[t]
        adv = delta + nt * gamma * lam * adv
        returns[t] = adv + values[t]
    a = returns - values
    return returns, (a - a.mean()) / (a.std() + 1e-8)


class PPOOracle:
    def __init__(self, ac, **kw):
        self.ac, self.a = ac, {**ARGS, **kw}
        self.opt = torch.optim.Adam(ac.parameters(), lr=self.a["learning_rate"])
        self.adapt_opt = torch.optim.Adam(ac.parameters(), lr=self.a["adaptation_module_learning_rate"])
        self.lr = self.a["learning_rate"]

    def update(self, hist, priv, actions, values, returns, advantages, old_logp, old_mu, old_sigma, perm):
        """All inputs flattened to [T*N, .]; perm = the minibatch permutation (same for every epoch)."""
        a, ac = self.a, self.ac
        mb = hist.shape[0] // a["num_mini_batches"]
        acc = [0.0, 0.0, 0.0, 0.0]
        for epoch in range(a["num_learning_epochs"]):
            for i in range(a["num_mini_batches"]):
                idx = perm[i * mb:(i + 1) * mb]
                h, p = hist[idx], priv[idx]
                d = ac.dist(h)
                logp = d.log_prob(actions[idx]).sum(-1)
                v = ac.value(h, p)
                mu, sigma, ent = d.mean, d.stddev, d.entropy().sum(-1)
                with torch.no_grad():
                    kl = torch.sum(torch.log(sigma / old_sigma[idx] + 1.e-5) + (old_sigma[idx] ** 2 + (old_mu[idx] - mu) ** 2) / (2.0 * sigma ** 2) - 0.5, -1).mean()
                    if kl > a["desired_kl"] * 2.0:
                        self.lr = max(1e-5, self.lr / 1.5)
                    elif kl < a["desired_kl"] / 2.0 and kl > 0.0:
                        self.lr = min(1e-2, self.lr * 1.5)
                    for g in self.opt.param_groups:
                        g["lr"] = self.lr
                ratio = torch.exp(logp - old_logp[idx].squeeze())
                A = advantages[idx].squeeze()
                surr = torch.max(-A * ratio, -A * torch.clamp(ratio, 1.0 - a["clip_param"], 1.0 + a["clip_param"])).mean()
                vt, R = values[idx], returns[idx]
                vc = vt + (v - vt).clamp(-a["clip_param"], a["clip_param"])
                vloss = torch.max((v - R).pow(2), (vc - R).pow(2)).mean()
                loss = surr + a["value_loss_coef"] * vloss - a["entropy_coef"] * ent.mean()
                self.opt.zero_grad()
                loss.backward()
                nn.utils.clip_grad_norm_(ac.parameters(), a["max_grad_norm"])
                self.opt.step()
                acc[0] += vloss.item(); acc[1] += surr.item()
                ntr = int(h.shape[0] // 5 * 4)
                pred = ac.adaptation_module(h)
                l_tr = F.mse_loss(pred[:ntr], p[:ntr]); l_te = F.mse_loss(pred[ntr:], p[ntr:])
                self.adapt_opt.zero_grad()
                l_tr.backward()
                self.adapt_opt.step()
                acc[2] += l_tr.item(); acc[3] += l_te.item()
        n = a["num_learning_epochs"] * a["num_mini_batches"]
        return acc[0] / n, acc[1] / n, acc[2] / n, acc[3] / n
