set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 100 python tools_profile_adj.py 2000 > gpurun_out/adj_plain.log 2>&1; cat gpurun_out/adj_plain.log
timeout 100 python tools_profile_adj.py 10000 >> gpurun_out/adj_plain.log 2>&1; tail -1 gpurun_out/adj_plain.log
