#!/bin/bash
# bn_bwd_apply second generation (EAT_BN_APPLY=v2): kernel-level + model-level parity under it, headline bench v1 / v2,
# and two existing knobs measured on the current code: EAT_FORK_WGRAD=1, EAT_TMA_BNMAX=256
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
T0=$SECONDS
timeout 100 python -m pytest tests/test_gpu_bn_bwd.py -q -m gpu -k apply 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -10
EAT_BN_APPLY=v2 timeout 250 python -m pytest tests/test_gpu_bn_bwd.py tests/test_gpu_mn_train.py tests/test_gpu_dymn.py tests/test_gpu_train_step.py -q -m gpu 2>&1 | grep -E "^E|passed|failed|Error" | cut -c1-300 | head -30
echo "tests took $((SECONDS-T0)) s"
for v in "EAT_BN_APPLY=v1" "EAT_BN_APPLY=v2" "EAT_BN_APPLY=v2 EAT_FORK_WGRAD=1" "EAT_BN_APPLY=v2 EAT_TMA_BNMAX=256"; do
  tag=$(echo $v | tr ' =' '__')
  echo "== $v"
  env $v EAT_BENCH_KERNELS=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/apply2_$tag.json 2> gpurun_out/apply2_$tag.err
  python -c "
import json
d=json.load(open('gpurun_out/apply2_$tag.json'))
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['gpu_launches'])
"
  grep -E "eat_bn_bwd_apply|eat_bn_bwd_reduce|eat_pw_tma_fwd |eat_pw_tc_wgrad" gpurun_out/apply2_$tag.err
done
echo "total $((SECONDS-T0)) s"
