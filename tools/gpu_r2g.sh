#!/bin/bash
# Round 2, call G (1 GPU): CMEM region table + tolerance mode, two compute streams in the host pipeline.
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
for fl in 0 4 131072; do
timeout 200 python bench.py --workload cmem --extra-flags $fl > gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.err; echo "bench cmem $fl rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json').read().strip().splitlines()[-1])
    print(d['value'], d['e2e']['value'], d['parity'], d['failed'])
except Exception as e:
    print('no line', e)
PY
tail -2 gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.err
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_march_cmem python bench.py --workload cmem --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_cmem.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_march_cmem_fast python bench.py --workload cmem --extra-flags 4 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_cmem_fast.log 2>&1
timeout 120 python tools/ab_quick.py --steps 4 --e2e --flags "0,32768" > gpurun_out/${tag}_ab_e2e.json 2> gpurun_out/${tag}_ab_e2e.err
python -c "
import json
d=json.load(open('gpurun_out/${tag}_ab_e2e.json'))
for k,v in d.items():
    if 'flags_' in k: print(k, v)
"
timeout 150 python bench.py > gpurun_out/${tag}_bench_orbit_1gpu.json 2> gpurun_out/${tag}_bench_orbit_1gpu.err; echo "bench orbit rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_orbit_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['parity'], d['failed'], d['tolerance_mode'], d['roofline']['frac'], d['roofline']['bound'])
PY
for w in hires256 hires512x4k; do
timeout 150 python bench.py --workload $w > gpurun_out/${tag}_bench_${w}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_1gpu.err; echo "bench $w rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_${w}_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['parity'], d['failed'], (d['tolerance_mode'] or {}).get('value'))
PY
done
