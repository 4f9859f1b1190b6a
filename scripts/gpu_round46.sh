#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python scripts/smoke_dbg.py 2>&1 | tail -12
echo "== exact fp32 GEMMs"; GEMM=simt timeout 300 python scripts/smoke_dbg.py 2>&1 | tail -4
