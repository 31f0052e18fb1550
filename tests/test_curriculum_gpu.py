"""Device-resident command curriculum (csrc/curriculum.cu, go1_curriculum_resample) against its host twin
LeggedRobot._resample_commands_host, which tests/test_resample_host.py pins to the reference's own
_resample_commands (legged_robot.py:710-824) -- so parity is transitive and bit-exact: float32 commands, int bins /
categories, float64 curriculum weights, the MT19937 words and position of every curriculum's RandomState, the category
generator's state.  Second test: a whole env stepped with the curriculum on the device vs on the host."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "walk-these-ways_b200", "compat"))


def _env(n, device_curriculum, **cmd_overrides):
    for m in [k for k in sys.modules if k.startswith("go1_gym.envs.base.legged_robot_config")]:
        del sys.modules[m]
    from go1_gym.envs.base.legged_robot_config import Cfg
    from go1_b200.train_config import apply_train_config
    from go1_gym.envs.go1.velocity_tracking import VelocityTrackingEasyEnv
    from go1_gym.envs.base.legged_robot import LeggedRobot
    apply_train_config(Cfg)
    Cfg.env.num_envs = n
    for k, v in cmd_overrides.items():
        setattr(Cfg.commands, k, v)
    LeggedRobot.device_curriculum = device_curriculum
    torch.manual_seed(0)            # creation-time domain randomisation draws from torch's global generators
    torch.cuda.manual_seed_all(0)
    try:
        env = VelocityTrackingEasyEnv(sim_device="cuda:0", headless=True, cfg=Cfg)
        env._device_curriculum()
    finally:
        LeggedRobot.device_curriculum = True
    return env


def _device_state(dc):
    torch.cuda.synchronize()
    return dict(weights=dc.weights.cpu().numpy(), mt=dc.mt.cpu().numpy().view(np.uint32), cat=int(dc.cat_rng.cpu().numpy().view(np.uint64)[0]),
                bins=dc.env_bins.cpu().numpy(), cats=dc.env_categories.cpu().numpy(), bins_f32=dc.env_bins_f32.cpu().numpy())


def _check_state(env, dc, tag, bins_snapshot):
    d = _device_state(dc)
    for i, cur in enumerate(env.curricula):
        assert np.array_equal(d["weights"][i], cur.weights), (tag, "weights", i)
        _, key, pos, _, _ = cur.rng.get_state()
        assert np.array_equal(d["mt"][i, :624], key) and int(d["mt"][i, 624]) == pos, (tag, "RandomState", i)
    assert d["cat"] == env._cat_rng.state, (tag, "category rng")
    assert np.array_equal(d["bins"], env.env_command_bins), (tag, "bins")
    assert np.array_equal(d["cats"], env.env_command_categories), (tag, "categories")
    assert np.array_equal(d["bins_f32"], bins_snapshot), (tag, "env_bins extras")


@pytest.mark.parametrize("mode", ["gaitwise", "nominal_binary", "exclusive", "balanced"])
def test_sequential_path_matches_host_twin(mode):
    """The same bit-exactness test with go1_curriculum_set_grouped(0): every call takes the category-by-category path (the default,
    grouped, path sends calls with <= 256 events through the category-parallel code)."""
    from go1_b200 import capi
    capi.lib().go1_curriculum_set_grouped(0)
    try:
        test_device_curriculum_matches_host_twin(mode)
    finally:
        capi.lib().go1_curriculum_set_grouped(1)


@pytest.mark.parametrize("mode", ["gaitwise", "nominal_binary", "exclusive", "balanced"])
def test_device_curriculum_matches_host_twin(mode):
    over = {"gaitwise": {}, "nominal_binary": dict(gaitwise_curricula=False, binary_phases=True),
            "exclusive": dict(gaitwise_curricula=False, exclusive_phase_offset=True),
            "balanced": dict(gaitwise_curricula=False, exclusive_phase_offset=False, balance_gait_distribution=True)}[mode]
    N = 512
    env = _env(N, True, **over)
    dc, core = env._dev_cur, env.core
    assert dc is not None
    dc.to_device()
    w0 = sum(float(c.weights.sum()) for c in env.curricula)
    snap = env.env_command_bins[:env.num_train_envs].astype(np.float32)
    rs = np.random.RandomState(11)
    ep_len, cols, thr = env._resample_constants()
    counts = [N, 0, 1, 5, 40, 3, 300, 7, 2, 9, 64, 1, 5, 480, 6]
    for r, k in enumerate(counts):
        which = r % 2 if r > 0 else 0
        ids = rs.choice(N, k, replace=False).astype(np.int64)             # event order is arbitrary (atomic slots)
        sums = np.zeros((k, 4), dtype=np.float32)
        for j in range(4):
            t = float(thr[cols.index(j)]) * float(ep_len) if j in cols else 1.0
            sums[:, j] = (t * (1.0 + rs.uniform(-0.2, 0.25 + 0.05 * r, size=k))).astype(np.float32)
        ev = np.zeros((k, 6), dtype=np.float32)
        ev[:, 0], ev[:, 1:5] = ids, sums
        core.events[which, :k] = torch.from_numpy(ev).cuda()
        core.event_count.zero_(); core.event_count[which] = k
        before = core.env("commands").clone()
        dc.resample(which)
        order = np.argsort(ids)
        want = env._resample_commands_host(ids[order], sums[order])       # advances the host objects
        torch.cuda.synchronize()
        sid = torch.from_numpy(ids[order]).cuda()
        if which == 0:
            assert int(dc.out_count.item()) == k
            assert np.array_equal(dc.out_ids[:k].cpu().numpy(), ids[order])
            got = dc.out_commands[:k].cpu().numpy()
            assert torch.equal(core.env("commands"), before)              # list 0 leaves the state to the reset kernel
        else:
            got = core.env("commands")[:, sid].t().cpu().numpy()
            assert float(core.env("command_sums")[:, sid].abs().sum()) == 0.0
        assert got.dtype == np.float32 and np.array_equal(got, want), (mode, r, k, np.abs(got - want).max() if k else 0)
        if which == 0 and k > 0:      # extras["env_bins"] is a snapshot taken by a step in which some env reset
            snap = env.env_command_bins[:env.num_train_envs].astype(np.float32)
        _check_state(env, dc, (mode, r, k), snap)
    # the rounds exercised successes (the weights grew); k = 512 / 480 / 300 pull > 624 MT words in one call (twist path)
    assert sum(float(c.weights.sum()) for c in env.curricula) > w0 + 1.0


def test_env_rollout_device_curriculum_equals_host_curriculum():
    """Two envs from the same seeds, one with the curriculum on the host (event list D2H + numpy), one on the device:
    every observation, reward, reset flag and command must agree bit for bit over a rollout with falls, time-outs and
    periodic resamples."""
    N, T = 256, 70
    envs = [_env(N, False), _env(N, True)]
    assert envs[0]._dev_cur is None and envs[1]._dev_cur is not None
    g = torch.Generator().manual_seed(3)
    ep0 = torch.randint(0, 1001, (N,), generator=g)
    for e in envs:
        e.reset()
        e.episode_length_buf = ep0.clone()
    n_reset = n_timeout = 0
    for t in range(T):
        a = (torch.randn(N, 12, generator=g) * (2.5 if t % 7 else 6.0)).cuda()
        outs = [e.step(a.clone()) for e in envs]
        torch.cuda.synchronize()
        for name, x, y in (("obs", outs[0][0], outs[1][0]), ("rew", outs[0][1], outs[1][1]), ("reset", outs[0][2], outs[1][2])):
            assert torch.equal(x, y), (t, name, float((x.float() - y.float()).abs().max()))
        assert torch.equal(envs[0].core.env("commands"), envs[1].core.env("commands")), t
        assert torch.equal(outs[0][3]["env_bins"], outs[1][3]["env_bins"]), t
        assert torch.equal(outs[0][3]["time_outs"], outs[1][3]["time_outs"]), t
        n_reset += int(outs[0][2].sum()); n_timeout += int(envs[0].core.timeout_u8.sum())
    assert n_reset > 20 and n_timeout > 3
    envs[1]._curriculum_to_host()
    for c0, c1 in zip(envs[0].curricula, envs[1].curricula):
        assert np.array_equal(c0.weights, c1.weights)
        assert np.array_equal(c0.rng.get_state()[1], c1.rng.get_state()[1]) and c0.rng.get_state()[2] == c1.rng.get_state()[2]
    assert np.array_equal(envs[0].env_command_bins, envs[1].env_command_bins)
    assert np.array_equal(envs[0].env_command_categories, envs[1].env_command_categories)
    ep = [dict(e.extras["train/episode"]) for e in envs]
    assert ep[0].keys() == ep[1].keys()
    for k in ep[0]:
        assert abs(float(ep[0][k]) - float(ep[1][k])) <= 1e-6 * max(1.0, abs(float(ep[0][k]))), k


@pytest.mark.parametrize("mode", ["gaitwise", "nominal_binary"])
def test_cross_rank_replay_equals_one_process_over_all_envs(mode):
    """two "ranks" of N envs whose event records are gathered (here: concatenated by hand; in production one NCCL
    all-gather per env step) and replayed by go1_curriculum_resample in cross-rank mode must evolve exactly like ONE process that
    owns 2N envs: identical curriculum weights, RandomState words and category stream on both ranks and on the reference, and every
    env gets the command / bin / category the single process gives to env rank * N + i."""
    from go1_b200.curriculum_dev import DeviceCurriculum
    from go1_gym.envs.base.legged_robot import _LOCAL_RANGE, _TASK_KEYS
    from go1_b200 import capi
    over = {"gaitwise": {}, "nominal_binary": dict(gaitwise_curricula=False, binary_phases=True)}[mode]
    N, W = 192, 2
    ref = _env(W * N, True, **over)
    ranks = [_env(N, True, **over) for _ in range(W)]
    for r, e in enumerate(ranks):
        e._dev_cur = DeviceCurriculum(e, _LOCAL_RANGE, _TASK_KEYS, emulate_world_rank=(W, r))
    dcs = [e._dev_cur for e in ranks]
    ref._dev_cur.to_device()
    for dc in dcs:
        dc.to_device()
    rs = np.random.RandomState(5)
    ep_len, cols, thr = ref._resample_constants()
    counts = [2 * N, 0, 3, 11, 40, 1, 300, 7, 64, 350, 5]          # > 256 and > 512 records exercise the global-scratch path
    for rnd, k in enumerate(counts):
        which = rnd % 2 if rnd > 0 else 0
        gids = rs.choice(W * N, k, replace=False).astype(np.int64)
        sums = np.zeros((k, 4), dtype=np.float32)
        for j in range(4):
            t = float(thr[cols.index(j)]) * float(ep_len) if j in cols else 1.0
            sums[:, j] = (t * (1.0 + rs.uniform(-0.2, 0.25 + 0.08 * rnd, size=k))).astype(np.float32)

        def load(core, ids, sm):
            ev = np.zeros((len(ids), 6), dtype=np.float32)
            ev[:, 0], ev[:, 1:5] = ids, sm
            core.events[which, :len(ids)] = torch.from_numpy(ev).cuda()
            core.event_count.zero_(); core.event_count[which] = len(ids)
        load(ref.core, gids, sums)
        ref._dev_cur.resample(which)
        for r, (e, dc) in enumerate(zip(ranks, dcs)):
            m = (gids >= r * N) & (gids < (r + 1) * N)
            load(e.core, gids[m] - r * N, sums[m])
            capi.check(e.core.L.go1_curriculum_pack(e.core._handle, dc._cfg_ref, dc._bufs_ref, capi.stream_ptr()), "pack")
        gathered = torch.stack([dc.xr_send for dc in dcs])                 # what all_gather_into_tensor delivers to every rank
        for dc in dcs:
            dc.xr_recv.copy_(gathered)
            dc.resample(which)
        torch.cuda.synchronize()
        R = _device_state(ref._dev_cur)
        for r, (e, dc) in enumerate(zip(ranks, dcs)):
            d = _device_state(dc)
            tag = (mode, rnd, k, r)
            assert np.array_equal(d["weights"], R["weights"]) and np.array_equal(d["mt"], R["mt"]) and d["cat"] == R["cat"], tag
            assert np.array_equal(d["bins"], R["bins"][r * N:(r + 1) * N]) and np.array_equal(d["cats"], R["cats"][r * N:(r + 1) * N]), tag
            sl = slice(r * N, (r + 1) * N)
            if which == 1:
                assert torch.equal(e.core.env("commands"), ref.core.env("commands")[:, sl]), tag
            else:
                kr = int(ref._dev_cur.out_count.item())
                rid = ref._dev_cur.out_ids[:kr].cpu().numpy(); rcmd = ref._dev_cur.out_commands[:kr].cpu().numpy()
                m = (rid >= r * N) & (rid < (r + 1) * N)
                kl = int(dc.out_count.item())
                assert kl == int(m.sum()), tag
                assert np.array_equal(dc.out_ids[:kl].cpu().numpy(), rid[m] - r * N) and np.array_equal(dc.out_commands[:kl].cpu().numpy(), rcmd[m]), tag
    assert float(R["weights"].sum()) > sum(float(c.weights.sum()) for c in ref.curricula) - 1e-9      # weights grew on the device
