#!/usr/bin/env python
"""Turns the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.
  python profiles/summarize.py launches gpurun_out/launches_r1_fp32.csv profiles/r1_launches_fp32.md "title"
  python profiles/summarize.py kernel   gpurun_out/prof_step_r1.ncu-rep  profiles/r1_step_kernel_ncu.md
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        n = r["Kernel Name"].split("(")[0].replace("void ", "")
        n = n[-70:]
        tot[n] += float(r["Metric Value"]); cnt[n] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over the whole command; {len(rows)} launches, "
                f"{T / 1e6:.1f} ms of kernel time (cold-cache, serialised: compare SHARES).\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, v in tot.most_common(25):
            f.write(f"| `{k}` | {cnt[k]} | {v / 1e6:.2f} | {100 * v / T:.1f}% |\n")


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full: {src.split('/')[-1]}\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']][:90]}  (launch id {r[idx['ID']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in keys:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\nWarp stall reasons (warps stalled per issue-active cycle):\n\n| reason | ratio |\n|---|---:|\n")
            st = sorted(((float(r[idx[h]] or 0), h) for h in stalls), reverse=True)
            for v, h in st[:8]:
                f.write(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |\n")
            f.write("\n")


def traffic(dst, pairs):
    """profiles/ncu_traffic.json: for each (kernel-name substring, .ncu-rep) pair the DRAM bytes per launch (mean over the captured
    launches of that kernel) that bench.py reports as roofline.traffic, plus pipe utilisation figures of the same launches."""
    import json
    out = {}
    try:
        out = json.load(open(dst))
    except Exception:
        pass
    for key, src in pairs:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        sel = [r for r in rows[2:] if key in r[idx["Kernel Name"]]]
        if not sel:
            continue

        def col(name, scale_units=True):
            vals = []
            for r in sel:
                v = float(r[idx[name]].replace(",", "") or 0)
                u = units[idx[name]]
                if scale_units:
                    v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
                vals.append(v)
            return vals
        rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
        dur = col("gpu__time_duration.sum", False)
        e = {"dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(sel), "launches_captured": len(sel), "source": src.split("/")[-1],
             "mean_duration_" + units[idx["gpu__time_duration.sum"]]: sum(dur) / len(dur)}
        for name, short in (("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
                            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
                            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")):
            if name in idx:
                v = col(name, False)
                w = [a * b for a, b in zip(v, dur)]
                e[short] = sum(w) / sum(dur)          # duration-weighted
        out[key] = e
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":         # traffic profiles/ncu_traffic.json key=file.ncu-rep ...
        traffic(sys.argv[2], [a.split("=", 1) for a in sys.argv[3:]])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else sys.argv[2])
    else:
        kernel(sys.argv[2], sys.argv[3])
