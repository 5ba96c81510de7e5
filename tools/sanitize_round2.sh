#!/bin/bash
# compute-sanitizer over the round-2 kernels (TMEM-parked twiddles, FP64 MAC / S5, single-launch polynomial multiply)
# on small batches; summaries into gpurun_out/r2_sanitizer.log
out=gpurun_out/r2_sanitizer.log
: > $out
for tool in memcheck racecheck synccheck; do
  for tgt in "ntt 0 300" "keyswitch 0 6" "polymul 1 300"; do
    echo "== $tool :: prof_target.py $tgt" >> $out
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/prof_target.py $tgt 2>&1 | grep -v "^=========     " | tail -6 >> $out
  done
done
cat $out
