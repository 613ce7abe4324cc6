#!/bin/bash
# A/B of the attention kernel variants on one box (parity first, then CUDA-event timings)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity (SWP default)"; timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -5
echo "== bench SWP=1 (recompute)"; timeout 120 python tools/attn_bench.py; timeout 120 python tools/attn_bench.py
echo "== bench SWP=0"; IVIT_ATTN_SWP=0 timeout 120 python tools/attn_bench.py; IVIT_ATTN_SWP=0 timeout 120 python tools/attn_bench.py
if [ -f i-vit_b200/csrc/libivit_b200_keepE.so ]; then
echo "== bench SWP=1 keep-E"; IVIT_B200_SO=$PWD/i-vit_b200/csrc/libivit_b200_keepE.so timeout 120 python tools/attn_bench.py
fi
} > gpurun_out/exp_attn.log 2>&1
cat gpurun_out/exp_attn.log
