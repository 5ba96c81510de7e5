python bench.py --impl reference > gpurun_out/r2_bench_v11_ref.json 2> gpurun_out/r2_bench_v11_ref.err
python bench.py > gpurun_out/r2_bench_v11.json 2> gpurun_out/r2_bench_v11.err
tail -2 gpurun_out/r2_bench_v11.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v11.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_launches_v11_bench.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_v11.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2_bench_v11_ref.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"ks",d["keyswitch"]["value"],"ks_sh",d["keyswitch_sharded"]["value"])
print("roofline",d["roofline"]["frac"],d["roofline"]["launch_s"],d["roofline"]["call_s"],d["roofline"]["inverse_kernel"])
print("e2e",d["e2e"]["value"],d["e2e"]["pageable"]["value"],d["e2e"]["pageable_registered"]["value"],d["e2e"]["copy_ceiling"]["GBps_per_gpu_each_way"])
print("ref",r["value"],r["cpu_baseline"]["cores"])
PY
