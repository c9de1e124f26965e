#!/bin/bash
# Round 2, call D: L1::no_allocate record loads as the default: A/B again, ncu, bench lines, whole GPU suite.
tag=${1:-r2d}
mkdir -p gpurun_out
show() { python -c "
import json,sys
d=json.load(open('$1'))
for k,v in d.items():
    if k.startswith('flags_'): print('$2', k, v['ms_min'], v['frames_per_s'], v['max_rel_diff_vs_first'])
"; }
AB="0,65536,n0,32768:fast_variant=0,32768:fast_variant=1,32768:fast_variant=2,32768:fast_variant=3,32768:fast_variant=4,32768:fast_kstep=0,32768:tile_swizzle=1,0:tile_swizzle=1,32768:tile_w_log2=3,32768:tile_w_log2=1,32768:tile_w_log2=3:tile_swizzle=1"
timeout 200 python tools/ab_quick.py --steps 4 --flags "$AB" > gpurun_out/${tag}_ab.json 2> gpurun_out/${tag}_ab.err
show gpurun_out/${tag}_ab.json default
for spec in "nnoalloc=n0" "mb6=32768,32768:fast_variant=4" "mb5=32768" "cmb6=0,65536" "l2_256=0,32768"; do
m=${spec%%=*}; fl=${spec#*=}
BIRP_B200_LIB=$PWD/tools/scratch/libbirp_b200_$m.so timeout 120 python tools/ab_quick.py --steps 4 --flags "$fl" > gpurun_out/${tag}_ab_$m.json 2> gpurun_out/${tag}_ab_$m.err
show gpurun_out/${tag}_ab_$m.json $m
done
timeout 150 ncu --set full --clock-control none --import-source on -k regex:march_cube_fast -s 2 -c 1 -f -o gpurun_out/${tag}_march_orbit_fast python bench.py --extra-flags 32768 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_orbit_fast.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:march_cube -s 2 -c 1 -f -o gpurun_out/${tag}_march_orbit python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_orbit.log 2>&1
for w in orbit hires256 hires512x4k; do
for fl in 0 32768; do
timeout 150 python bench.py --workload $w --extra-flags $fl > gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.err; echo "bench $w $fl rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/${tag}_bench_${w}_flags${fl}_1gpu.json'))
print(d['value'], d['e2e']['value'], d['parity'], d['failed'])"
done; done
timeout 500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${tag}_pytest.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
