set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
python tools_bench_shapes.py 64 16 32 128 256 > gpurun_out/shapes.log 2>&1
cat gpurun_out/shapes.log
