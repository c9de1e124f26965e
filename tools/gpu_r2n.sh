#!/bin/bash
# Round 2, call N (1 GPU): warps per CTA chosen at launch (cta_warps): new parity tests, A/B by knob on large images.
tag=${1:-r2n}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta or large_image or interleaved or every_kernel_variant or config" > gpurun_out/${tag}_pytest_subset.log 2>&1
tail -2 gpurun_out/${tag}_pytest_subset.log
for w in 0 8 4 2; do
  timeout 100 python tools/ab_rows.py --bands 128 --warps $w > gpurun_out/${tag}_ab_rows_512x4k_w$w.json 2>> gpurun_out/${tag}_ab.err
  timeout 100 python tools/ab_rows.py --bands 128 --warps $w --flags 32768 > gpurun_out/${tag}_ab_rows_512x4k_fast_w$w.json 2>> gpurun_out/${tag}_ab.err
  timeout 100 python tools/ab_rows.py --bands 128 --warps $w --cube 256 --image 1024 --nsamples 2048 > gpurun_out/${tag}_ab_rows_256x1k_w$w.json 2>> gpurun_out/${tag}_ab.err
  timeout 100 python tools/ab_rows.py --bands 128 --warps $w --cube 256 --image 1024 --nsamples 2048 --flags 32768 > gpurun_out/${tag}_ab_rows_256x1k_fast_w$w.json 2>> gpurun_out/${tag}_ab.err
  timeout 100 python tools/ab_rows.py --bands 64 --warps $w --cube 256 --image 512 > gpurun_out/${tag}_ab_rows_256x512_w$w.json 2>> gpurun_out/${tag}_ab.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${tag}_ab_rows_*.json')):
    try:
        d=json.load(open(f)); print(f.split('ab_rows_')[1], d['whole']['ms_min'], [ (k,v['ms_min'],v['shares_equal_whole']) for k,v in d.items() if k.startswith('band_')])
    except Exception as e: print(f, e)
PY
tail -3 gpurun_out/${tag}_ab.err
for w in hires256 hires512x4k; do
  timeout 150 python bench.py --workload $w > gpurun_out/${tag}_bench_${w}_1gpu.json 2> gpurun_out/${tag}_bench_${w}_1gpu.err; echo "bench $w rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench_${w}_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'], d['failed'], (d['tolerance_mode'] or {}).get('value'))
PY
done
