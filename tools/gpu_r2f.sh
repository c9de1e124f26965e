#!/bin/bash
# Round 2, call F (1 GPU): suite with marked oct records / cfitsio-format FITS, realistic-cube workloads, CMEM source profile.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
for w in orbit-zeros orbit-neg cmem; do
timeout 200 python bench.py --workload $w > gpurun_out/${tag}_bench_${w}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_1gpu.err; echo "bench $w rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_${w}_1gpu.json').read().strip().splitlines()[-1])
    print(d['value'], d['e2e']['value'], d['parity'], d['failed'], (d.get('tolerance_mode') or {}).get('value'))
except Exception as e:
    print('no line', e)
PY
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_march_cmem python bench.py --workload cmem --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_cmem.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:march_cube -s 2 -c 1 -f -o gpurun_out/${tag}_march_orbit_zeros python bench.py --workload orbit-zeros --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_orbit_zeros.log 2>&1
timeout 120 python tools/load_time.py > gpurun_out/${tag}_load_time.json 2> gpurun_out/${tag}_load_time.err
cat gpurun_out/${tag}_load_time.json
timeout 120 python tools/ab_quick.py --steps 4 --e2e --flags "0,32768" > gpurun_out/${tag}_ab_e2e.json 2> gpurun_out/${tag}_ab_e2e.err
python -c "
import json
d=json.load(open('gpurun_out/${tag}_ab_e2e.json'))
for k,v in d.items():
    if 'flags_' in k: print(k, v)
"
