#!/bin/bash
# TMA residual tile of the residual-block GEMMs (IVIT_GEMM_RTMA=1 default / 0): parity, then the step time of the three configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity"; timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_swin_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
echo "== IVIT_GEMM_RTMA=$v"; IVIT_GEMM_RTMA=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['configs'].items()})"
done
} > gpurun_out/exp_gemm.log 2>&1
cat gpurun_out/exp_gemm.log
