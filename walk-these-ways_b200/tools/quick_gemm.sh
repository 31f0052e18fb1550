#!/bin/bash
# Inner loop of the GEMM work: learner parity tests + one bench run with the per-launch GEMM timing CSV.  $1 = tag, rest = env assignments
TAG=${1:-q}; shift
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_ppo_gpu.py tests/test_misc_kernels_gpu.py -q -x 2>&1 | tail -15 > $O/${TAG}_tests.txt
env "$@" GO1_GEMM_TIMING_CSV=$O/${TAG}_gemm_launches.csv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -3 $O/${TAG}_tests.txt; tail -c 600 $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
