#!/bin/bash
# full ncu capture of one plain NTT kernel launch with per-phase op mix:  bash tools/ncu_ntt_kernel.sh k_ntt_inv 2
K=$1; SKIP=${2:-2}
mkdir -p /tmp/ncu gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:"$K" -s $SKIP -c 1 -o /tmp/ncu/nx python tools/prof_target.py ntt 1 4096 > /tmp/ncu/e.log 2>&1
python tools/ncu_summary.py /tmp/ncu/nx.ncu-rep 2>&1 | grep -v "fp64\.\(max\|min\|sum\)\|ops_path\|cycles_active\.\(max\|min\|sum\)\|realtime" > gpurun_out/r2_ncu_${K}_v11_summary.txt
(python tools/ncu_opmix.py /tmp/ncu/nx.ncu-rep $K; python tools/ncu_hot.py /tmp/ncu/nx.ncu-rep $K 0.006) > gpurun_out/r2_ncu_${K}_v11_opmix_phases.txt 2>&1
tail -1 /tmp/ncu/e.log
