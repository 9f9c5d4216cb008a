timeout 400 python -m pytest tests/test_gpu_graph.py -q -x -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_pf.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])"
tail -3 gpurun_out/bench_pf.err
