#!/bin/bash
# Round 2, call K (1 GPU): the closing validation of the committed binary -- GPU suite, smoke(), compute-sanitizer over
# every kernel variant, ncu re-captures (traffic.json must describe this binary), launch list, every bench line.
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitizer_smoke.py > gpurun_out/${tag}_sanitizer_${tool}.txt 2>&1
  echo "sanitizer $tool rc=$?"; tail -2 gpurun_out/${tag}_sanitizer_${tool}.txt
done
for spec in "orbit:0" "orbit:32768" "hires256:0" "hires512x4k:0" "cmem:0"; do
  w=${spec%%:*}; fl=${spec#*:}
  k=march_cube; [ $w = cmem ] && k=march_kernel
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/${tag}_march_${w}_flags${fl} python bench.py --workload $w --extra-flags $fl --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_${w}_flags${fl}.log 2>&1
  echo "ncu $w $fl rc=$?"
done
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_orbit.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_launches.log 2>&1
for spec in "orbit:0" "orbit:32768" "orbit-zeros:0" "orbit-neg:0" "orbit-nearest:0" "hires256:0" "hires512x4k:0" "cmem:0" "cmem:4"; do
  w=${spec%%:*}; fl=${spec#*:}
  timeout 200 python bench.py --workload $w --extra-flags $fl > gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.err; echo "bench $w $fl rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'].get('bit_exact'), d['failed'], d['roofline']['bound'], d['roofline']['frac'], (d.get('tolerance_mode') or {}).get('value'), d['clocks'])
except Exception as e:
    print('no line', e)
PY
done
timeout 200 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_1gpu.json 2> gpurun_out/${tag}_bench_reference_1gpu.err; echo "reference arm rc=$?"
head -c 200 gpurun_out/${tag}_bench_reference_1gpu.json; echo
timeout 200 python tools/load_time.py > gpurun_out/${tag}_load_time.json 2> gpurun_out/${tag}_load_time.err; cat gpurun_out/${tag}_load_time.json | head -c 600; echo
timeout 120 python tools/ab_quick.py --steps 4 --e2e --flags "0,32768" > gpurun_out/${tag}_ab_e2e.json 2> gpurun_out/${tag}_ab_e2e.err
python -c "
import json
d=json.load(open('gpurun_out/${tag}_ab_e2e.json'))
for k,v in d.items():
    if 'flags_' in k: print(k, v)
"
