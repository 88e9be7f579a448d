set -x
mkdir -p gpurun_out
timeout 150 python tools_bench_hunt.py all 512 64 > gpurun_out/hunt_plain.log 2>&1
cat gpurun_out/hunt_plain.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hunt4.csv python tools_bench_hunt.py cfg4 512 16 > gpurun_out/ncu_hunt4.log 2>&1
wc -l gpurun_out/launches_hunt4.csv
