set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
python tools_bench_shapes.py 256 64 > gpurun_out/shapes.log 2>&1
cat gpurun_out/shapes.log
ncu --set full --clock-control none --import-source on -k regex:tile_ -s 20 -c 5 -o gpurun_out/prof_tile256 python tools_profile_tile.py OrbitKS 256 256 64 3 256 > gpurun_out/ncu_tile256.log 2>&1
