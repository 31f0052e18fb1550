O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2t_tests.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --breakdown > $O/r2t_bench.json 2> $O/r2t_bench.err
tail -4 $O/r2t_tests.txt; python -c "
import json;d=json.loads(open('$O/r2t_bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['phase_ms'],d['losses'],d['roofline_sim_step']['kernel_ms'],d['roofline']['frac'])"; tail -3 $O/r2t_bench.err
