# One GPU-box session: tests, bench, launch list.   gpurun --timeout 900 -- 'bash tools/gpu_check.sh'
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -c 2500 gpurun_out/bench.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-hunt --no-cpu-baseline --no-side > gpurun_out/ncu_bench.log 2>&1
