#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/bench_dw.py --batch 256 2>&1 | tail -1
