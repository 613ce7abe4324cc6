#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py -x -q -k "embed or engine or full_batch or graph" 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"embed_tokens|patchify" -c 4 python tools/profile_forward.py --batch 256 --forwards 2 2>&1 | grep -E "embed_tokens|patchify|gpu__time" | head -12
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], d['parity'])"
} > gpurun_out/exp_embed.log 2>&1
cat gpurun_out/exp_embed.log
