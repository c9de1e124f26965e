#!/bin/bash
# Round 2, call L (1 GPU): the closing validation of the committed binary -- GPU suite, smoke(), compute-sanitizer over
# every kernel variant, every bench line, launch list, and ncu re-captures summarised ON the box (the .ncu-rep files
# stay in /tmp there: five of them exceed what gpurun copies back).
tag=${1:-r2l}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1
tail -1 gpurun_out/${tag}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitizer_smoke.py > gpurun_out/${tag}_sanitizer_${tool}.txt 2>&1
  echo "sanitizer $tool rc=$?"; tail -1 gpurun_out/${tag}_sanitizer_${tool}.txt
done
for spec in "orbit:0" "orbit:32768" "orbit-zeros:0" "orbit-neg:0" "orbit-nearest:0" "hires256:0" "hires512x4k:0" "cmem:0" "cmem:4"; do
  w=${spec%%:*}; fl=${spec#*:}
  timeout 200 python bench.py --workload $w --extra-flags $fl > gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.err; echo "bench $w $fl rc=$?"
done
timeout 200 python bench.py --impl reference > gpurun_out/${tag}_bench_reference_1gpu.json 2> gpurun_out/${tag}_bench_reference_1gpu.err; echo "reference arm rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_orbit.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_launches.log 2>&1
for spec in "orbit:0" "orbit:32768" "orbit-zeros:0" "hires256:0" "hires512x4k:0" "cmem:0" "cmem:4"; do
  w=${spec%%:*}; fl=${spec#*:}
  k=march_cube; [ $w = cmem ] && k=march_kernel
  key=$w; name=${w}; [ $fl != 0 ] && key="${w}+flags${fl}" && name="${w}_flags${fl}"
  rep=/tmp/${tag}_march_${name}.ncu-rep
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o ${rep%.ncu-rep} python bench.py --workload $w --extra-flags $fl --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_${name}.log 2>&1
  echo "ncu $w $fl rc=$?"
  python tools/ncu_traffic.py $rep gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.json "$key" profiles/${tag}_march_${name}.md "closing capture of round 2 (committed binary), bench.py --workload $w --extra-flags $fl" > /dev/null 2> gpurun_out/${tag}_traffic_${name}.err
  ncu -i $rep --page raw --csv > gpurun_out/${tag}_march_${name}_raw.csv 2>/dev/null
  cp profiles/${tag}_march_${name}.md gpurun_out/ 2>/dev/null
done
cp profiles/traffic.json gpurun_out/${tag}_traffic.json
cp /tmp/${tag}_march_orbit.ncu-rep gpurun_out/ 2>/dev/null
timeout 200 python tools/load_time.py > gpurun_out/${tag}_load_time.json 2> gpurun_out/${tag}_load_time.err
du -sh gpurun_out
