#!/bin/bash
# full ncu capture of one keyswitch stage kernel with per-phase op mix:  bash tools/ncu_ks_kernel.sh k_ks_intt1 [skip]
K=$1; SKIP=${2:-0}
mkdir -p /tmp/ncu gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:"$K" -s $SKIP -c 1 -o /tmp/ncu/ksx python tools/prof_target.py keyswitch 1 ${KS_ITEMS:-444} > /tmp/ncu/e.log 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ksx.ncu-rep $K; python tools/ncu_hot.py /tmp/ncu/ksx.ncu-rep $K 0.006) > gpurun_out/r2_ncu_${K}_v11_opmix_phases.txt 2>&1
tail -1 /tmp/ncu/e.log
