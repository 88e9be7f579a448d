set -x
mkdir -p gpurun_out
python tools_profile_tile.py OrbitKS 256 256 64 20 256 > gpurun_out/tile256_plain.log 2>&1
python tools_profile_tile.py RelativeOrbitKS 64 64 4096 20 256 > gpurun_out/tile64_plain.log 2>&1
cat gpurun_out/tile256_plain.log gpurun_out/tile64_plain.log
ncu --set full --clock-control none --import-source on -k regex:tile_ -s 24 -c 6 -o gpurun_out/prof_tile256 python tools_profile_tile.py OrbitKS 256 256 64 3 256 > gpurun_out/ncu_tile256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_ -s 24 -c 6 -o gpurun_out/prof_tile64 python tools_profile_tile.py RelativeOrbitKS 64 64 4096 3 256 > gpurun_out/ncu_tile64.log 2>&1
tail -3 gpurun_out/ncu_tile256.log gpurun_out/ncu_tile64.log
