#!/bin/bash
# round 2, final pass: full ncu captures of the plain 52-bit forward / inverse kernels and of the keyswitch
# stage kernels as they are at the end of the round, summarised on the GPU box -> gpurun_out/r2_ncu_*_v5*.txt
set -u
TAG=${1:-v5}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 280 $NCU -k regex:"k_ntt_fwd|k_ntt_inv" -s 3 -c 3 -o /tmp/ncu/ntt64 python tools/prof_target.py ntt 1 4096 > /tmp/ncu/a.log 2>&1
python tools/ncu_summary.py /tmp/ncu/ntt64.ncu-rep > gpurun_out/r2_ncu_ntt64_${TAG}_summary.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ntt64.ncu-rep k_ntt_fwd; python tools/ncu_hot.py /tmp/ncu/ntt64.ncu-rep k_ntt_fwd 0.006) > gpurun_out/r2_ncu_ntt64_fwd_${TAG}_opmix_phases.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ntt64.ncu-rep k_ntt_inv; python tools/ncu_hot.py /tmp/ncu/ntt64.ncu-rep k_ntt_inv 0.006) > gpurun_out/r2_ncu_ntt64_inv_${TAG}_opmix_phases.txt 2>&1
timeout 400 $NCU -k regex:"k_ks_" -s 6 -c 6 -o /tmp/ncu/ks python tools/prof_target.py keyswitch 1 ${KS_ITEMS:-444} > /tmp/ncu/d.log 2>&1
python tools/ncu_summary.py /tmp/ncu/ks.ncu-rep > gpurun_out/r2_ncu_ks_${TAG}_summary.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ks.ncu-rep k_ks_ntt2f; python tools/ncu_hot.py /tmp/ncu/ks.ncu-rep k_ks_ntt2f 0.008) > gpurun_out/r2_ncu_ks_ntt2f_${TAG}_opmix_phases.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ks.ncu-rep k_ks_mac_fp64; python tools/ncu_hot.py /tmp/ncu/ks.ncu-rep k_ks_mac_fp64 0.01) > gpurun_out/r2_ncu_ks_mac_${TAG}_opmix_phases.txt 2>&1
tail -2 /tmp/ncu/a.log /tmp/ncu/d.log
