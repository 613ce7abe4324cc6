#!/bin/bash
# A/B of the attention kernel variants on one box (parity first, then CUDA-event timings)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity (default build)"; timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" 2>&1 | tail -5
echo "== bench OW=1"; timeout 120 python tools/attn_bench.py; timeout 120 python tools/attn_bench.py
echo "== bench OW=0"; IVIT_ATTN_OW=0 timeout 120 python tools/attn_bench.py; IVIT_ATTN_OW=0 timeout 120 python tools/attn_bench.py
echo "== bench DeiT-S shape OW=1/0"; NSEQ=128 HEADS=6 timeout 120 python tools/attn_bench.py; IVIT_ATTN_OW=0 NSEQ=128 HEADS=6 timeout 120 python tools/attn_bench.py
} > gpurun_out/exp_attn.log 2>&1
cat gpurun_out/exp_attn.log
