set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 150 python tools_bench_hunt.py all 512 64 > gpurun_out/hunt_plain.log 2>&1
cat gpurun_out/hunt_plain.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hunt4.csv python tools_bench_hunt.py cfg4 512 16 > gpurun_out/ncu_hunt4.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_hunt3.csv python tools_bench_hunt.py cfg3 256 16 > gpurun_out/ncu_hunt3.log 2>&1
wc -l gpurun_out/launches_hunt4.csv gpurun_out/launches_hunt3.csv
