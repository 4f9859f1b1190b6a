#!/bin/bash
# round 2: full GPU test-suite + default bench with the TMA pointwise kernel in the engine
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
EAT_TEST_REPORT=gpurun_out/parity_report.txt timeout 900 python -m pytest tests -m gpu -q -rs 2>&1 | grep -v "^\s*$" | tail -60 > gpurun_out/full1_pytest.log
tail -3 gpurun_out/full1_pytest.log
EAT_BENCH_KERNELS=1 python bench.py --steps 10 --warmup 3 > gpurun_out/full1_bench.json 2> gpurun_out/full1_bench.err
cat gpurun_out/full1_bench.err | tail -30
python - <<'PY'
import json
d = json.load(open("gpurun_out/full1_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], {k: v.get("value") for k, v in d["gpu_baseline"].items()})
PY
