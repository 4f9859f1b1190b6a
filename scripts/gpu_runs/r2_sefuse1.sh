#!/bin/bash
# SE blocks: one-pass squeeze-excitation + BatchNorm-backward reduce (EAT_SE_FUSED=1) and the second-generation pooling
# kernel (EAT_POOL=v2): new kernel-level tests, train-step parity under the flags, the f4 script tests, then the headline
# bench with the flags on (flags-off reference: profiles/r02_bench_fp32_b256.json, 34.3 ms/step on the same code)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -m pytest tests/test_gpu_bn_bwd.py tests/test_gpu_f4.py -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -30
echo "== v2 / fused"
EAT_POOL=v2 EAT_SE_FUSED=1 timeout 200 python -m pytest tests/test_gpu_bn_bwd.py tests/test_gpu_mn_train.py tests/test_gpu_mn.py tests/test_gpu_train_step.py -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -30
echo "tests took $((SECONDS-T0)) s"
for v in "EAT_SE_FUSED=1 EAT_POOL=v1" "EAT_SE_FUSED=1 EAT_POOL=v2"; do
  tag=$(echo $v | tr ' =' '__')
  echo "== $v"
  env $v EAT_BENCH_KERNELS=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/sefuse1_$tag.json 2> gpurun_out/sefuse1_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/sefuse1_$tag.json'))
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['gpu_launches'])
"
  grep -E "eat_se_|eat_bn_act_pool|eat_bn_bwd_reduce|eat_bn_bwd_apply" gpurun_out/sefuse1_$tag.err
done
echo "total $((SECONDS-T0)) s"
