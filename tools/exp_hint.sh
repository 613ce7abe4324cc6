#!/bin/bash
# A/B: mbarrier try_wait with / without the suspend-time hint (two builds of the same library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NOH=$PWD/i-vit_b200/csrc/libivit_b200_nohint.so
{
echo "== parity (hinted build)"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_swin_gpu.py -x -q 2>&1 | tail -3
for rep in 1 2; do
echo "== hint"; timeout 120 python tools/attn_bench.py; timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], [ (s['name'], round(s['ms']*1000,1)) for s in d['roofline']['per_shape']], {k:v['ms_per_step'] for k,v in d['configs'].items()})"
echo "== no hint"; IVIT_B200_SO=$NOH timeout 120 python tools/attn_bench.py; IVIT_B200_SO=$NOH timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], [ (s['name'], round(s['ms']*1000,1)) for s in d['roofline']['per_shape']], {k:v['ms_per_step'] for k,v in d['configs'].items()})"
done
} > gpurun_out/exp_hint.log 2>&1
cat gpurun_out/exp_hint.log
