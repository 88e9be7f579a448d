set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hunt > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hunt > gpurun_out/ncu1.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
cat gpurun_out/bench.log gpurun_out/bench_ref.log
