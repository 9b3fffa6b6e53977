#!/bin/bash
# round 2, call B: 3 dequant warpgroups + 1 CTA/SM; full GPU tests; sweeps; timelines; ncu source-level profile
mkdir -p gpurun_out
echo "=== hang probe"; timeout 120 python -u scripts/gpu_hang_probe.py 4096 4096 32 2>&1 | tail -4
timeout 120 python -u scripts/gpu_hang_probe.py 6144 4096 1 2>&1 | tail -4
echo "=== layer chain"; timeout 300 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -4
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | grep -v "^DEBUG\|^$" | tail -40
echo "=== shapes"; timeout 300 python -u scripts/gpu_int4_layer.py shapes 2>&1 | tail -14
echo "=== sweep"
for e in "AO_B200_TS_CTAS_PER_SM=2" "AO_B200_TS_MIN_UNITS=2" "AO_B200_TS_MIN_UNITS=8" "AO_B200_TS_PREFETCH=0" "AO_B200_NO_PDL=1"; do
  echo "--- $e"; env $e timeout 200 python -u scripts/gpu_int4_layer.py one fused 2>&1 | tail -2
done
echo "=== timeline"; for s in 6144x4096 4096x4096 28672x4096 4096x14336; do timeout 120 python -u scripts/gpu_timeline.py 1,32 $s 2>&1 | tail -8; done
echo "=== ncu"; timeout 600 ncu --set full --clock-control none --import-source on --sampling-interval 0 -k regex:ts_gemm -s 3 -c 1 -f -o gpurun_out/r02_gateup_m32 python scripts/gpu_ncu_one.py 32 28672 4096 5 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on --sampling-interval 0 -k regex:ts_gemm -s 3 -c 1 -f -o gpurun_out/r02_qkv_m32 python scripts/gpu_ncu_one.py 32 6144 4096 5 2>&1 | tail -3
