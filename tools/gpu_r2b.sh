#!/bin/bash
# Round 2, call B: the tolerance-mode march (BIRP_FAST_MARCH): GPU tests, A/B of its variants, ncu capture, bench line.
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast or whole_orbit" --durations=5 > gpurun_out/${tag}_pytest_fast.log 2>&1
tail -5 gpurun_out/${tag}_pytest_fast.log
AB="0,65536,32768:fast_variant=0,32768:fast_variant=1,32768:fast_variant=2,32768:fast_variant=3,32768:fast_variant=4,32768:fast_variant=0:fast_kstep=0,32768:fast_variant=2:fast_kstep=0,32768:fast_variant=4:fast_kstep=0,32768:fast_variant=2:tile_swizzle=1,32768:fast_variant=0:tile_swizzle=1,0:tile_swizzle=1,32768:fast_variant=2:tile_w_log2=3,32768:fast_variant=2:tile_w_log2=1,32768:fast_variant=2:tile_w_log2=3:tile_swizzle=1,32768:fast_variant=2:tile_w_log2=1:tile_swizzle=1"
timeout 200 python tools/ab_quick.py --steps 4 --flags "$AB" > gpurun_out/${tag}_ab_fast.json 2> gpurun_out/${tag}_ab_fast.err
cat gpurun_out/${tag}_ab_fast.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    if k.startswith('flags_'): print(k, v['ms_min'], v['frames_per_s'], v['max_rel_diff_vs_first'])
"
for mb in 6 5; do
BIRP_B200_LIB=$PWD/tools/scratch/libbirp_b200_mb$mb.so timeout 120 python tools/ab_quick.py --steps 4 --flags "0,32768:fast_variant=0,32768:fast_variant=2,32768:fast_variant=4,32768:fast_variant=2:tile_swizzle=1" > gpurun_out/${tag}_ab_fast_mb$mb.json 2> gpurun_out/${tag}_ab_fast_mb$mb.err
python -c "
import json,sys
d=json.load(open('gpurun_out/${tag}_ab_fast_mb$mb.json'))
for k,v in d.items():
    if k.startswith('flags_'): print('mb$mb', k, v['ms_min'], v['frames_per_s'], v['max_rel_diff_vs_first'])
"
done
timeout 150 python bench.py --extra-flags 32768 > gpurun_out/${tag}_bench_orbit_fast_1gpu.json 2> gpurun_out/${tag}_bench_orbit_fast_1gpu.err; echo "bench rc=$?"
head -c 300 gpurun_out/${tag}_bench_orbit_fast_1gpu.json; echo; tail -3 gpurun_out/${tag}_bench_orbit_fast_1gpu.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:march_cube_fast -s 2 -c 1 -f -o gpurun_out/${tag}_march_orbit_fast python bench.py --extra-flags 32768 --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_orbit_fast.log 2>&1
ls -la gpurun_out | grep ${tag}
