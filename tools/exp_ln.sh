#!/bin/bash
# LayerNorm: four-rows-per-warp kernel (default) against the 16-lane kernel (IVIT_LN_VARIANT=9); parity first
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
for v in 0 9 0 9; do
echo "== IVIT_LN_VARIANT=$v"; IVIT_LN_VARIANT=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['configs'].items()})"
done
} > gpurun_out/exp_ln.log 2>&1
cat gpurun_out/exp_ln.log
