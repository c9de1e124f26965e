#!/bin/bash
# Round 2, call S (1 GPU, the last minutes of the budget): CMEM after the above-the-magnetopause exact path and the bare
# MUFU.RSQ estimates -- whole GPU suite first, then the CMEM bench lines with a fresh ncu capture (each line printed
# again after its capture, so that its roofline uses the new instruction count), memcheck last.
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/${tag}_pytest.log 2>&1
tail -1 gpurun_out/${tag}_pytest.log
for fl in 0 4; do
  name=cmem; key=cmem; [ $fl != 0 ] && key="cmem+flags${fl}" && name="cmem_flags${fl}"
  rep=/tmp/${tag}_march_${name}.ncu-rep
  timeout 100 python bench.py --workload cmem --extra-flags $fl > gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.err; echo "bench cmem $fl rc=$?"
  timeout 100 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 2 -c 1 -f -o ${rep%.ncu-rep} python bench.py --workload cmem --extra-flags $fl --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_${name}.log 2>&1
  python tools/ncu_traffic.py $rep gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json "$key" profiles/${tag}_march_${name}.md "CMEM march after the last changes of round 2, bench.py --workload cmem --extra-flags $fl" > /dev/null 2> gpurun_out/${tag}_traffic_${name}.err
  ncu -i $rep --page raw --csv > gpurun_out/${tag}_march_${name}_raw.csv 2>/dev/null
  cp profiles/${tag}_march_${name}.md gpurun_out/ 2>/dev/null
  cp profiles/traffic.json gpurun_out/${tag}_traffic.json
  timeout 100 python bench.py --workload cmem --extra-flags $fl > gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.err; echo "bench cmem $fl rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_cmem_flags${fl}_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'].get('pixels_within_1e-5'), d['failed'], d['roofline']['frac'])
PY
  if [ $fl = 0 ]; then
    timeout 200 compute-sanitizer --tool memcheck python tools/sanitizer_smoke.py > gpurun_out/${tag}_sanitizer_memcheck.txt 2>&1
    echo "sanitizer memcheck rc=$?"; tail -1 gpurun_out/${tag}_sanitizer_memcheck.txt
  fi
done
timeout 60 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
