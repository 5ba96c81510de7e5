#!/bin/bash
# Full ncu captures of the hot kernels, summarised on the GPU box (the reports themselves are too
# large to bring back): writes gpurun_out/r1_ncu_*_v7*.txt.   Usage: bash tools/ncu_round.sh
set -u
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 280 $NCU -k regex:"k_ntt_fwd|k_ntt_inv" -s 2 -c 4 -o /tmp/ncu/ntt64 python tools/prof_target.py ntt 1 4096 > /tmp/ncu/a.log 2>&1
timeout 280 $NCU -k regex:k_ntt_small -s 2 -c 2 -o /tmp/ncu/ntt32 python tools/prof_target.py ntt28 1 4096 > /tmp/ncu/b.log 2>&1
timeout 200 $NCU -k regex:'k_dyadic$' -s 1 -c 1 -o /tmp/ncu/dyadic python tools/prof_target.py dyadic > /tmp/ncu/c.log 2>&1
timeout 400 $NCU -k regex:"k_ks_" -s 7 -c 7 -o /tmp/ncu/ks python tools/prof_target.py keyswitch 1 114 > /tmp/ncu/d.log 2>&1
for r in ntt64 ntt32 dyadic ks; do
  python tools/ncu_summary.py /tmp/ncu/$r.ncu-rep > gpurun_out/r1_ncu_${r}_v7_summary.txt 2>&1
done
(python tools/ncu_opmix.py /tmp/ncu/ntt64.ncu-rep k_ntt_fwd; python tools/ncu_hot.py /tmp/ncu/ntt64.ncu-rep k_ntt_fwd 0.006) > gpurun_out/r1_ncu_ntt64_fwd_v7_opmix_phases.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ntt64.ncu-rep k_ntt_inv; python tools/ncu_hot.py /tmp/ncu/ntt64.ncu-rep k_ntt_inv 0.006) > gpurun_out/r1_ncu_ntt64_inv_v7_opmix_phases.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ntt32.ncu-rep k_ntt_small; python tools/ncu_hot.py /tmp/ncu/ntt32.ncu-rep k_ntt_small 0.006) > gpurun_out/r1_ncu_ntt32_fwd_v7_opmix_phases.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/dyadic.ncu-rep k_dyadic) > gpurun_out/r1_ncu_dyadic_v7_opmix.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ks.ncu-rep k_ks_mac_fast; python tools/ncu_hot.py /tmp/ncu/ks.ncu-rep k_ks_mac_fast 0.01) > gpurun_out/r1_ncu_ks_mac_v7_opmix_phases.txt 2>&1
ls -la gpurun_out
