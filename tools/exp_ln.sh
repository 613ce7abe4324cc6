#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity"; timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm" 2>&1 | tail -3
echo "== r4"; timeout 120 python tools/rowops_bench.py 2>&1 | grep layernorm; timeout 120 python tools/rowops_bench.py 2>&1 | grep layernorm
echo "== 16-lane"; IVIT_LN_VARIANT=9 timeout 120 python tools/rowops_bench.py 2>&1 | grep layernorm; IVIT_LN_VARIANT=9 timeout 120 python tools/rowops_bench.py 2>&1 | grep layernorm
timeout 300 ncu --set full --clock-control none --import-source on -k regex:layernorm --launch-skip 3 -c 1 -f -o gpurun_out/ln_r4b python tools/rowops_bench.py 2>&1 | tail -1
} > gpurun_out/exp_ln.log 2>&1
cat gpurun_out/exp_ln.log
