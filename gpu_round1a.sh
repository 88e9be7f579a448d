set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ks_warp32 -s 10 -c 2 -o gpurun_out/prof python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.log
