#!/bin/bash
# last pass of the round: memcheck + synccheck (+ racecheck on the keyswitch) over the kernels as shipped
out=gpurun_out/r2_sanitizer_final.log
: > $out
for tool in memcheck synccheck; do
  for tgt in "ntt 0 300" "keyswitch 0 6" "polymul 1 300"; do
    echo "== $tool :: prof_target.py $tgt" >> $out
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/prof_target.py $tgt 2>&1 | grep -v "^=========     " | tail -3 >> $out
  done
done
echo "== racecheck :: prof_target.py keyswitch 0 6" >> $out
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python tools/prof_target.py keyswitch 0 6 2>&1 | grep -v "^=========     " | tail -3 >> $out
cat $out
python -c "import __graft_entry__ as g; g.smoke()"
