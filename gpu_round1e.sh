set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in 3 4; do python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-hunt --variant $v > gpurun_out/bench_v$v.log 2>&1; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hunt --variant 4 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ks_warp32 -s 10 -c 1 -o gpurun_out/prof_v4 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hunt --variant 4 > gpurun_out/ncu_v4.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
grep -o '"value": [0-9.]*' gpurun_out/bench_v*.log
