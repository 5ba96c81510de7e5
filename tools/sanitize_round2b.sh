#!/bin/bash
# racecheck: the default forward kernel (buffer reuse ordered by a shared-memory atomic + the TMA mbarrier, which the
# tool does not model) against the same kernel with warp_tail=0 (block barrier in front of the prefetch)
out=gpurun_out/r2_racecheck_warp_tail.log
: > $out
for opts in "warp_tail=1" "warp_tail=0"; do
  for tgt in "ntt 0 300" "polymul 1 300"; do
    echo "== racecheck :: HB_OPTS=$opts prof_target.py $tgt" >> $out
    HB_OPTS=$opts timeout 900 compute-sanitizer --tool racecheck --print-limit 1 python tools/prof_target.py $tgt 2>&1 | grep -v "^=========     " | tail -4 >> $out
  done
done
cat $out
python tools/quick_time.py > gpurun_out/quick_time_fence.json 2>&1; tail -c 1500 gpurun_out/quick_time_fence.json
