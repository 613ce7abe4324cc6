#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attention_pipe_kernel|gemm_i8_tcgen05_kernel" --launch-skip 6 -c 5 -f -o gpurun_out/hot_r2b python tools/profile_forward.py --batch 256 --forwards 1 2>&1 | tail -2
} > gpurun_out/exp_prof.log 2>&1
cat gpurun_out/exp_prof.log; ls -la gpurun_out/*.ncu-rep
