#!/bin/bash
# stride-2 depthwise data gradient with the expand BatchNorm's backward reduce in its epilogue (EAT_DGRAD_BNRED=1):
# kernel-level + train-step parity, then the headline bench without / with it (both with the SE-fused reduce, pool v2)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -m pytest tests/test_gpu_bn_bwd.py tests/test_gpu_mn_train.py -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -30
timeout 100 python -m pytest tests/test_gpu_f4.py -q -m gpu -k "native or golden_logits" 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -10
echo "tests took $((SECONDS-T0)) s"
for v in "EAT_DGRAD_BNRED=0" "EAT_DGRAD_BNRED=1"; do
  tag=$(echo $v | tr ' =' '__')
  echo "== $v"
  env $v EAT_BENCH_KERNELS=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/dgred1_$tag.json 2> gpurun_out/dgred1_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/dgred1_$tag.json'))
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['gpu_launches'])
"
  grep -E "eat_se_bn|eat_bn_act_pool|eat_bn_bwd_reduce|eat_dw_conv_dgrad" gpurun_out/dgred1_$tag.err
done
echo "total $((SECONDS-T0)) s"
