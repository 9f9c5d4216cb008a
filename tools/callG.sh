start=$(date +%s)
timeout 900 python -m pytest tests/ -q -m gpu -x > gpurun_out/pytest_full5.txt 2>&1
echo "pytest rc=$? secs=$(( $(date +%s) - start ))"
tail -8 gpurun_out/pytest_full5.txt
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_stem.json 2> gpurun_out/bench_stem.err
echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_stem.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['per_kind'])"
