#!/bin/bash
timeout 300 python -m pytest tests/ -q -m gpu --timeout 120 2>&1 | tail -3
echo "=== lowp bench"; timeout 200 python -u scripts/gpu_probe_lowp.py --stage bench 2>&1 | grep -E "M=32|M= 1 gate|M= 1 down|M= 1 q" | head -27
