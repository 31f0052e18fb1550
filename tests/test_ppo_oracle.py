"""Pins oracle/ppo_oracle.py to the vectors produced by the reference's own ppo_cse code (tests/golden/ppo.npz)."""
import os

import numpy as np
import torch

from oracle.ppo_oracle import ActorCriticOracle, PPOOracle, gae
from ppo_golden_util import seeded_weights, sample_tensor

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_reference_ppo_cycle():
    torch.set_num_threads(4)
    g = np.load(os.path.join(HERE, "golden", "ppo.npz"))
    ac = ActorCriticOracle()
    w = seeded_weights({k: tuple(v.shape) for k, v in ac.state_dict().items()})
    ac.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    T_ = lambda k: torch.from_numpy(np.ascontiguousarray(g[k]))
    hist, priv, eps = T_("in/hist"), T_("in/priv"), T_("in/eps")
    T, N = hist.shape[:2]
    with torch.no_grad():
        acts, vals, logps, mus = [], [], [], []
        for t in range(T):
            d = ac.dist(hist[t])
            a = d.mean + d.stddev * eps[t]
            acts.append(a); vals.append(ac.value(hist[t], priv[t])); logps.append(d.log_prob(a).sum(-1, keepdim=True)); mus.append(d.mean)
        actions, values, logp, mu = torch.stack(acts), torch.stack(vals), torch.stack(logps), torch.stack(mus)
        last_v = ac.value(T_("last/hist"), T_("last/priv"))
        returns, adv = gae(T_("in/rew").unsqueeze(-1), T_("in/done").unsqueeze(-1), values, last_v)
    for name, got in (("actions", actions), ("values", values), ("actions_log_prob", logp), ("mu", mu), ("returns", returns), ("advantages", adv)):
        assert np.allclose(got.numpy(), g[f"storage/{name}"], rtol=1e-4, atol=2e-5), name
    f = lambda x: x.flatten(0, 1)
    ppo = PPOOracle(ac)
    vl, sl, al, atl = ppo.update(f(hist), f(priv), f(actions), f(values), f(returns), f(adv), f(logp), f(mu), torch.ones_like(f(mu)),
                                 torch.from_numpy(g["in/perm"]))
    ref = g["update/losses"]
    assert abs(vl - ref[0]) < 1e-3 * abs(ref[0]) and abs(sl - ref[1]) < 1e-3 and abs(al - ref[2]) < 1e-3 * abs(ref[2]) and abs(atl - ref[5]) < 1e-3 * abs(ref[5])
    assert abs(ppo.lr - float(g["update/learning_rate"])) < 1e-12
    for k, v in ac.state_dict().items():
        got, want = sample_tensor(v.numpy()), g[f"final/{k}"]
        assert np.allclose(got[:-2], want[:-2], atol=2e-4), k
