#!/bin/bash
# Round 2, call M (1 GPU): CTA size A/B (8 / 4 / 2 warps) on the orbit sweep and on one rank's share of config 4.
tag=${1:-r2m}
mkdir -p gpurun_out
for v in default w4 w2; do
  lib=birp_cpp_b200/libbirp_b200.so; [ $v != default ] && lib=tools/scratch/libbirp_b200_$v.so
  BIRP_B200_LIB=$PWD/$lib timeout 200 python tools/ab_rows.py --bands 128,32,8 > gpurun_out/${tag}_ab_rows_$v.json 2> gpurun_out/${tag}_ab_rows_$v.err
  cat gpurun_out/${tag}_ab_rows_$v.json; tail -2 gpurun_out/${tag}_ab_rows_$v.err
  BIRP_B200_LIB=$PWD/$lib timeout 200 python tools/ab_rows.py --bands 128 --flags 32768 > gpurun_out/${tag}_ab_rows_fast_$v.json 2> gpurun_out/${tag}_ab_rows_fast_$v.err
  cat gpurun_out/${tag}_ab_rows_fast_$v.json; tail -2 gpurun_out/${tag}_ab_rows_fast_$v.err
  BIRP_B200_LIB=$PWD/$lib timeout 120 python tools/ab_quick.py --steps 4 --flags "0,32768,n0" > gpurun_out/${tag}_ab_quick_$v.json 2> gpurun_out/${tag}_ab_quick_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/${tag}_ab_quick_$v.json'))
for k,v in d.items():
    if 'flags_' in k: print('$v', k, v['ms_min'], v['same_as_first'])
"
done
