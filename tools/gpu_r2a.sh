#!/bin/bash
# Round 2, call A: GPU suite with the lean march as default, sanitizer over every kernel, bench line with the
# issue roofline + exit-code gate, ncu captures of the HBM-resident configs.
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
timeout 120 python bench.py > gpurun_out/${tag}_bench_orbit_1gpu.json 2> gpurun_out/${tag}_bench_orbit_1gpu.err; echo "bench rc=$?"
head -c 400 gpurun_out/${tag}_bench_orbit_1gpu.json; echo
for tool in memcheck racecheck initcheck; do
  timeout 420 compute-sanitizer --tool $tool python tools/sanitizer_smoke.py > gpurun_out/${tag}_sanitizer_${tool}.txt 2>&1
  tail -3 gpurun_out/${tag}_sanitizer_${tool}.txt
done
for w in hires256 hires512x4k; do
  timeout 150 ncu --set full --clock-control none --import-source on -k regex:march_cube -s 2 -c 1 -f -o gpurun_out/${tag}_march_${w} python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_${w}.log 2>&1
  timeout 100 python bench.py --workload $w > gpurun_out/${tag}_bench_${w}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_1gpu.err; echo "bench $w rc=$?"
  head -c 300 gpurun_out/${tag}_bench_${w}_1gpu.json; echo
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:march_cube -s 2 -c 1 -f -o gpurun_out/${tag}_march_orbit python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_orbit.log 2>&1
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_orbit.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_launches.log 2>&1
ls -la gpurun_out | tail -30
