#!/usr/bin/env python
"""Host-side wall-clock split of the 24-step rollout (perf_counter around the phases of one env step; no profiler)."""
import os
import sys
import time
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402  (puts the package on sys.path)
import torch  # noqa: E402

acc, cnt = defaultdict(float), defaultdict(int)


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def timed(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[label] += time.perf_counter() - t
            cnt[label] += 1
    setattr(obj, name, timed)


def main():
    env, runner = bench.build_training(4096, "cuda:0", 1)
    od = env.get_observations()
    st = [od["obs"], od["privileged_obs"], od["obs_history"]]

    def it():
        o, p, h, _ = runner.rollout(*st)
        st[:] = [o, p, h]
        with torch.inference_mode():
            runner.alg.compute_returns(h[:env.num_train_envs], p[:env.num_train_envs])
        runner.alg.update()
    for _ in range(3):
        it()
    torch.cuda.synchronize()
    base, core = env.env, env.env.core
    wrap(runner.alg, "act", "alg.act")
    wrap(runner.alg, "process_env_step", "alg.process_env_step")
    wrap(env, "step", "wrapper.step (total)")
    wrap(base, "_step_device", "  _step_device")
    wrap(base, "_apply_pending_interval_resample", "  interval resample")
    wrap(core, "step", "  core.step launch")
    wrap(core, "fetch_events", "  fetch_events (sync)")
    wrap(base, "_post_physics_step_callback_host", "  gravity callback")
    wrap(base, "reset_idx", "  reset_idx (total)")
    wrap(base, "_resample_commands_host", "    resample_commands_host")
    wrap(core, "reset_idx", "    core.reset_idx")
    wrap(core, "set_commands", "    core.set_commands")
    wrap(base, "_fill_extras", "    fill_extras")
    n = 4
    t0 = time.perf_counter()
    for _ in range(n):
        o, p, h, _ = runner.rollout(*st)
        st[:] = [o, p, h]
        torch.cuda.synchronize()
        runner.alg.storage.clear()
    tot = time.perf_counter() - t0
    steps = n * 24
    print(f"rollout wall {tot / n * 1e3:.2f} ms  ({tot / steps * 1e6:.0f} us/step)")
    for k, v in acc.items():
        print(f"{k:34s} {v / steps * 1e6:8.1f} us/step   calls/step {cnt[k] / steps:.2f}")
    # GPU side of the same loop: kernel time per env step (CUPTI)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            o, p, h, _ = runner.rollout(*st)
            st[:] = [o, p, h]
            torch.cuda.synchronize()
            runner.alg.storage.clear()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot_us = sum(e.device_time_total for e in rows)
    print(f"GPU busy {tot_us / 48:.0f} us/step")
    for e in rows[:14]:
        print(f"  {e.key[:70]:70s} {e.device_time_total / 48:8.1f} us/step  x{e.count / 48:.1f}")


if __name__ == "__main__":
    main()
