#!/bin/bash
# compute-sanitizer over the kernels added late in round 2 (four-rows-per-warp LayerNorm, TVM-semantics operators) and the
# LayerNorm variants they share arithmetic with
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py tests/test_tvm_mode_gpu.py -x -q -k "layernorm or tvm or softmax or gelu" -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/memcheck_r2b.log 2>&1
( timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_kernels_gpu.py -x -q -k "layernorm_i16_i8" -p no:cacheprovider 2>&1 | tail -6 ) > gpurun_out/racecheck_r2b.log 2>&1
cat gpurun_out/memcheck_r2b.log gpurun_out/racecheck_r2b.log
