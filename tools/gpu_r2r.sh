#!/bin/bash
# Round 2, call R (1 GPU): CMEM after the classify-then-evaluate filter and the above-the-magnetopause exact path: whole
# GPU suite, sanitizer, CMEM bench lines with a fresh ncu capture (each line is printed again after its capture, so that its
# roofline uses the new instruction count), orbit line.
tag=${1:-r2r}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1
tail -1 gpurun_out/${tag}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitizer_smoke.py > gpurun_out/${tag}_sanitizer_${tool}.txt 2>&1
  echo "sanitizer $tool rc=$?"; tail -1 gpurun_out/${tag}_sanitizer_${tool}.txt
done
for fl in 0 4; do
  timeout 200 python bench.py --workload cmem --extra-flags $fl > gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.err; echo "bench cmem $fl rc=$?"
  name=cmem; key=cmem; [ $fl != 0 ] && key="cmem+flags${fl}" && name="cmem_flags${fl}"
  rep=/tmp/${tag}_march_${name}.ncu-rep
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o ${rep%.ncu-rep} python bench.py --workload cmem --extra-flags $fl --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_${name}.log 2>&1
  python tools/ncu_traffic.py $rep gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json "$key" profiles/${tag}_march_${name}.md "CMEM march after the last changes of round 2, bench.py --workload cmem --extra-flags $fl" > /dev/null 2> gpurun_out/${tag}_traffic_${name}.err
  ncu -i $rep --page raw --csv > gpurun_out/${tag}_march_${name}_raw.csv 2>/dev/null
  cp profiles/${tag}_march_${name}.md gpurun_out/ 2>/dev/null
  timeout 200 python bench.py --workload cmem --extra-flags $fl > gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.err; echo "bench cmem $fl rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'].get('pixels_within_1e-5'), d['failed'], d['roofline']['frac'])
PY
done
cp profiles/traffic.json gpurun_out/${tag}_traffic.json
cp /tmp/${tag}_march_cmem.ncu-rep gpurun_out/ 2>/dev/null
timeout 100 python bench.py > gpurun_out/${tag}_bench_orbit_flags0_1gpu.json 2> gpurun_out/${tag}_bench_orbit_flags0_1gpu.err; echo "bench orbit rc=$?"
du -sh gpurun_out
