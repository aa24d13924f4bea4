mkdir -p gpurun_out/s17
python tools/host_profile.py 101 > gpurun_out/s17/host101.txt 2>&1
python tools/host_profile.py 1025 > gpurun_out/s17/host1025.txt 2>&1
echo finished
