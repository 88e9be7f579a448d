set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --variant 0 > gpurun_out/bench_v0.log 2>&1
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --variant 1 > gpurun_out/bench_v1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ks_warp32 -s 10 -c 1 -o gpurun_out/prof_v1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --variant 1 > gpurun_out/ncu_v1.log 2>&1
tail -n 5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
grep -o '"value": [0-9.]*' gpurun_out/bench_v0.log gpurun_out/bench_v1.log
