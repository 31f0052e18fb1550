O=gpurun_out
timeout 900 python -m pytest tests/test_sim_gpu.py tests/test_dr_gpu.py tests/test_curriculum_gpu.py tests/test_terrain_gpu.py tests/test_play_gpu.py -q -x 2>&1 | tail -4 > $O/r2u_tests.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown > $O/r2u_bench.json 2> $O/r2u_bench.err
tail -3 $O/r2u_tests.txt; python -c "
import json;d=json.loads(open('$O/r2u_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['phase_ms'],d['losses'],d['roofline_sim_step']['kernel_ms'])"; tail -3 $O/r2u_bench.err
