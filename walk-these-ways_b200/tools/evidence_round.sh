#!/bin/bash
# The round's evidence in one gpurun call: full GPU suite + smoke, bench lines of every BASELINE config, reference arm, torch-profiler kernel
# table, ncu launch list and `ncu --set full` captures of the fused step kernel, the tcgen05 GEMMs and the fused MLP tail.  $1 = tag.
TAG=${1:-r2}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/${TAG}_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline > $O/${TAG}_bench_n1_breakdown.json 2>> $O/${TAG}_bench.err
if [[ -z "$EVIDENCE_LIGHT" ]]; then
GO1_UPDATE_STREAMS=0 timeout 600 python bench.py --steps 5 --warmup 3 --breakdown --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_bench_n1_onestream.json 2>> $O/${TAG}_bench.err
fi
GO1_GEMM_TIMING_CSV=$O/${TAG}_gemm_launches.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --config rough_dr --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_rough_dr.json 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --config mob16k --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_mob16k.json 2>> $O/${TAG}_bench.err
timeout 900 python bench.py --config sweep --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_sweep.json 2>> $O/${TAG}_bench.err
( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 ) > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gemm-roofline --profile > /dev/null 2>> $O/${TAG}_bench.err
cp $O/kernels_torchprof.txt $O/${TAG}_kernels_torchprof.txt
timeout 600 python walk-these-ways_b200/tools/run_reference_scripts.py --iterations 2 --num-envs 4096 --out $O/${TAG}_reference_scripts.json > $O/${TAG}_reference_scripts.log 2>&1
if [[ -z "$EVIDENCE_LIGHT" ]]; then
timeout 200 python walk-these-ways_b200/tools/tail_bench.py > $O/${TAG}_tail_bench.txt 2>&1
timeout 200 python walk-these-ways_b200/tools/epi_bench.py > $O/${TAG}_epi_bench.txt 2>&1
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 1800 --csv --log-file $O/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:go1_step_kernel -s 30 -c 1 -f -o $O/${TAG}_step_kernel \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_ncu_step.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 300 -c 16 -f -o $O/${TAG}_gemm \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp_tail -s 60 -c 3 -f -o $O/${TAG}_tail \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-roofline > $O/${TAG}_ncu_tail.log 2>&1
du -sh $O; tail -3 $O/${TAG}_tests.txt; cat $O/${TAG}_smoke.txt | tail -1
