#!/bin/bash
# round 2, second pass: full ncu capture of the plain 52-bit forward / inverse kernels as they are now
# (forward: full correction every other stage), summarised on the GPU box -> gpurun_out/r2_ncu_ntt64_*.txt
set -u
TAG=${1:-v4}
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 280 $NCU -k regex:"k_ntt_fwd|k_ntt_inv" -s 4 -c 4 -o /tmp/ncu/ntt64 python tools/prof_target.py ntt 1 4096 > /tmp/ncu/a.log 2>&1
python tools/ncu_summary.py /tmp/ncu/ntt64.ncu-rep > gpurun_out/r2_ncu_ntt64_${TAG}_summary.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ntt64.ncu-rep k_ntt_fwd; python tools/ncu_hot.py /tmp/ncu/ntt64.ncu-rep k_ntt_fwd 0.006) > gpurun_out/r2_ncu_ntt64_fwd_${TAG}_opmix_phases.txt 2>&1
(python tools/ncu_opmix.py /tmp/ncu/ntt64.ncu-rep k_ntt_inv; python tools/ncu_hot.py /tmp/ncu/ntt64.ncu-rep k_ntt_inv 0.006) > gpurun_out/r2_ncu_ntt64_inv_${TAG}_opmix_phases.txt 2>&1
tail -3 /tmp/ncu/a.log
