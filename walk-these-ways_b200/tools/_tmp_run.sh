O=gpurun_out
timeout 600 python -m pytest tests/test_ppo_gpu.py -q -k "mn_major" 2>&1 | tail -15 > $O/r2n_tests.txt
for V in "GO1_TF32_WIDE_MINTILES=60" "GO1_TF32_WIDE_MINTILES=9"; do
  env $V timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gemm-roofline > $O/r2n.json 2>> $O/r2n.err
  python -c "
import json;d=json.loads(open('$O/r2n.json').read().strip().splitlines()[-1]);print('$V',d['ms_per_step'],d['losses'])"
done
tail -12 $O/r2n_tests.txt
