set -x
mkdir -p gpurun_out
python tools_profile_adj.py 2000 > gpurun_out/adj_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:x1_kernel -s 400 -c 1 -o gpurun_out/prof_adj python tools_profile_adj.py 600 > gpurun_out/ncu_adj.log 2>&1
cat gpurun_out/adj_plain.log
