set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -c 3000 gpurun_out/bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-hunt --no-cpu-baseline --no-side > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv
