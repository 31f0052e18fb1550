O=gpurun_out
timeout 900 python -m pytest tests/test_runner_gpu.py tests/test_curriculum_gpu.py tests/test_reference_scripts_gpu.py -q -x 2>&1 | tail -4 > $O/r2y_tests.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown --no-gemm-roofline > $O/r2y_bench.json 2> $O/r2y_bench.err
GO1_STEP_FORK=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown --no-gemm-roofline > $O/r2y_bench_nofork.json 2>> $O/r2y_bench.err
tail -3 $O/r2y_tests.txt; for f in r2y_bench r2y_bench_nofork; do python -c "
import json;d=json.loads(open('$O/$f.json').read().strip().splitlines()[-1]);print('$f',d['ms_per_step'],d['value'],d['phase_ms'],d['losses'])"; done; tail -3 $O/r2y_bench.err
