#!/bin/bash
# round 2: full ncu captures of the fused kernels (and the staged keyswitch for the traffic comparison),
# summarised on the GPU box -> gpurun_out/r2_ncu_*.txt
set -u
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 280 $NCU -k regex:k_polymul_fused -s 1 -c 1 -o /tmp/ncu/pm python tools/prof_target.py polymul 1 592 > /tmp/ncu/a.log 2>&1
timeout 280 $NCU -k regex:"k_ks_" -s 5 -c 5 -o /tmp/ncu/ksf python tools/prof_target.py keyswitch_fused 1 111 > /tmp/ncu/b.log 2>&1
for r in pm ksf; do
  python tools/ncu_summary.py /tmp/ncu/$r.ncu-rep > gpurun_out/r2_ncu_${r}_summary.txt 2>&1
done
(python tools/ncu_opmix.py /tmp/ncu/pm.ncu-rep k_polymul_fused; python tools/ncu_hot.py /tmp/ncu/pm.ncu-rep k_polymul_fused 0.01) > gpurun_out/r2_ncu_pm_opmix_phases.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ksf.ncu-rep k_ks_fused; python tools/ncu_hot.py /tmp/ncu/ksf.ncu-rep k_ks_fused 0.01) > gpurun_out/r2_ncu_ksf_opmix_phases.txt 2>&1
tail -3 /tmp/ncu/a.log /tmp/ncu/b.log
