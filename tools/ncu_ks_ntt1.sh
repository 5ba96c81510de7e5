#!/bin/bash
# full ncu capture of the keyswitch S2 kernel (k_ks_ntt1) with per-phase op mix
mkdir -p /tmp/ncu gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:"k_ks_ntt1" -s 1 -c 1 -o /tmp/ncu/ks1 python tools/prof_target.py keyswitch 1 ${KS_ITEMS:-444} > /tmp/ncu/e.log 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ks1.ncu-rep k_ks_ntt1; python tools/ncu_hot.py /tmp/ncu/ks1.ncu-rep k_ks_ntt1 0.006) > gpurun_out/r2_ncu_ks_ntt1_v9_opmix_phases.txt 2>&1
tail -2 /tmp/ncu/e.log
