#!/bin/bash
# One gpurun call's worth of evidence: GPU test suite, bench lines of the BASELINE configs, the reference arm, ncu launch list and
# `ncu --set full` captures of the step kernel and the dominant GEMM.  Everything lands in gpurun_out/ (tag = $1).
#   gpurun --timeout 1500 -- 'bash walk-these-ways_b200/tools/gpu_round.sh r2a'
TAG=${1:-r2}
WHAT=${2:-all}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
if [[ $WHAT == all || $WHAT == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/${TAG}_tests.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1
fi
if [[ $WHAT == *dist2* ]]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 \
      > $O/${TAG}_bench_flat_n2.json 2> $O/${TAG}_bench_flat_n2.err
  GO1_SHARED_CURRICULUM=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 \
      > $O/${TAG}_bench_flat_n2_percurr.json 2>> $O/${TAG}_bench_flat_n2.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --config rough_dr --steps 5 --warmup 3 \
      > $O/${TAG}_bench_rough_dr_n2.json 2>> $O/${TAG}_bench_flat_n2.err
fi
if [[ $WHAT == *gemmtune* ]]; then
  for V in ${GEMMTUNE_VARIANTS:-"GO1_X=0" "GO1_FUSE_BIAS_GRAD=0" "GO1_UPDATE_STREAMS=1"}; do
    echo "== $V" >> $O/${TAG}_gemmtune.txt
    env $V timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['ms_per_step'], d['roofline']['kernel_ms_per_iteration'], d['roofline']['frac'])
" >> $O/${TAG}_gemmtune.txt
  done
fi
if [[ $WHAT == *gemmcsv* ]]; then
  GO1_GEMM_TIMING_CSV=$O/${TAG}_gemm_launches.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_gemmcsv.json 2>&1
  GO1_TEST_GROUPED=1 timeout 300 python -m pytest tests/test_curriculum_gpu.py -q -k grouped 2>&1 | tail -5 > $O/${TAG}_grouped_test.txt
fi
if [[ $WHAT == *tailbench* ]]; then
  timeout 200 python walk-these-ways_b200/tools/tail_bench.py > $O/${TAG}_tail_bench.txt 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tail -s 4 -c 2 -f -o $O/${TAG}_tail python walk-these-ways_b200/tools/tail_bench.py > $O/${TAG}_ncu_tail.log 2>&1
fi
if [[ $WHAT == *onetest* ]]; then
  timeout ${ONETEST_TIMEOUT:-200} python -m pytest tests/test_ppo_gpu.py -q -x -k "${ONETEST:-fused}" 2>&1 | tail -40 > $O/${TAG}_onetest.txt
fi
if [[ $WHAT == *blocks* ]]; then
  GO1_SWEEP_BLOCKS=32,64,128 GO1_SWEEP_ENVS=4096,16384 timeout 300 python walk-these-ways_b200/tools/sim_sweep.py > $O/${TAG}_sim_blocks.txt 2>&1
fi
if [[ $WHAT == *hunt* ]]; then
  timeout 600 python walk-these-ways_b200/tools/nan_hunt.py --config rough_dr --envs 4096 --train 8 > $O/${TAG}_nan_hunt.txt 2>&1
fi
if [[ $WHAT == *refscripts* ]]; then
  timeout 900 python walk-these-ways_b200/tools/run_reference_scripts.py --iterations 2 --num-envs 4096 --out $O/${TAG}_reference_scripts.json > $O/${TAG}_reference_scripts.log 2>&1
fi
if [[ $WHAT == *compare* ]]; then
  timeout 900 python walk-these-ways_b200/tools/train_compare.py --iterations 200 --out $O/${TAG}_train_compare.json > $O/${TAG}_train_compare.log 2>&1
fi
if [[ $WHAT == all || $WHAT == *bench* ]]; then
  timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_flat.json 2> $O/${TAG}_bench_flat.err
  timeout 600 python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline > $O/${TAG}_bench_flat_breakdown.json 2>> $O/${TAG}_bench_flat.err
fi
if [[ $WHAT == all || $WHAT == *benchall* ]]; then
  timeout 600 python bench.py --config rough_dr --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_rough_dr.json 2> $O/${TAG}_bench_rough_dr.err
  timeout 600 python bench.py --config mob16k --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_mob16k.json 2> $O/${TAG}_bench_mob16k.err
fi
if [[ $WHAT == all || $WHAT == *sweep* ]]; then
  timeout 900 python bench.py --config sweep --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_sweep.json 2> $O/${TAG}_bench_sweep.err
fi
if [[ $WHAT == all || $WHAT == *ref* ]]; then
  ( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
fi
if [[ $WHAT == all || $WHAT == *ncu* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 2300 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_ncu_launches.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:go1_step_kernel -s 30 -c 2 -f -o $O/${TAG}_step_kernel \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_ncu_step.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 400 -c 12 -f -o $O/${TAG}_gemm \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_ncu_gemm.log 2>&1
fi
du -sh $O; ls -la $O | tail -30
