#!/bin/bash
# Round-end GPU check (one gpurun call): the whole GPU test suite, one ncu --set full capture of the dominant kernel,
# the bench line, the ncu launch list, then the other workloads as time allows.  Outputs under gpurun_out/<tag>_*.
tag=${1:-check}
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
timeout 90 ncu --set full --clock-control none --import-source on -k regex:march_cube -s 2 -c 1 -f -o gpurun_out/${tag}_march_orbit python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 90 python bench.py > gpurun_out/${tag}_bench_orbit_1gpu.json 2> gpurun_out/${tag}_bench_orbit_1gpu.err
head -c 300 gpurun_out/${tag}_bench_orbit_1gpu.json; echo
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench_orbit.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_launches.log 2>&1
for w in orbit-nearest hires256 hires512x4k; do
timeout 60 python bench.py --workload $w > gpurun_out/${tag}_bench_${w}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_1gpu.err
head -c 200 gpurun_out/${tag}_bench_${w}_1gpu.json; echo
done
