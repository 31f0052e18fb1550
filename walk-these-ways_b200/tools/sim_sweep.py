#!/usr/bin/env python
"""BASELINE config 5: env-count sweep of the fused sim-step kernel alone (CUDA events, 512 MiB L2 flush between launches).
Prints env-steps/s and the algorithmic HBM rate (3,328 B per env-step) against the measured HBM peak."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import torch  # noqa: E402


def main():
    from env_golden_util import train_sim_config
    from go1_b200.sim import SimCore
    hbm, _, src = bench.peaks()
    flush = torch.empty(512 * 1024 * 1024 // 4, device="cuda")
    rows = []
    from go1_b200 import capi
    blocks = [int(x) for x in os.environ.get("GO1_SWEEP_BLOCKS", "0").split(",")]
    counts = [int(x) for x in os.environ.get("GO1_SWEEP_ENVS", "1024,4096,16384,65536,131072,262144").split(",")]
    for n, blk in [(n, b) for n in counts for b in blocks]:
        capi.lib().go1_sim_set_step_block(blk)
        _, c, _ = train_sim_config(n)
        core = SimCore(c, device="cuda:0")
        core.env("root_pos")[2].fill_(0.34)
        actions = torch.zeros(n, 12, device="cuda")
        ts = []
        for i in range(9):
            flush.fill_(float(i))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); core.step(actions, common_step=100 + i, mode=0); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        gbs = bench.SIM_BYTES_PER_ENV_STEP * n / (ms * 1e-3) / 1e9
        rows.append({"envs": n, "block": blk, "kernel_ms": round(ms, 4), "env_steps_per_s": round(n / (ms * 1e-3)), "algorithmic_GBps": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / hbm, 4)})
        print(rows[-1], flush=True)
        del core
        torch.cuda.empty_cache()
    print(json.dumps({"hbm_peak_GBps": hbm, "peak_source": src, "rows": rows}))


if __name__ == "__main__":
    main()
