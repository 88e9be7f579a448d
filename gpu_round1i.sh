set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
python tools_profile_adj.py 2000 > gpurun_out/adj_plain.log 2>&1
python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/adj_plain.log
tail -c 600 gpurun_out/bench.log
