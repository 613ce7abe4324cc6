#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_swin_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
echo "== r4"; timeout 120 python tools/rowops_bench.py 2>&1 | grep layernorm; timeout 120 python tools/rowops_bench.py 2>&1 | grep layernorm
echo "== 16-lane"; IVIT_LN_VARIANT=9 timeout 120 python tools/rowops_bench.py 2>&1 | grep layernorm
echo "== step"; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['configs'].items()}, d['parity'])"
} > gpurun_out/exp_ln.log 2>&1
cat gpurun_out/exp_ln.log
