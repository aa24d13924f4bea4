mkdir -p gpurun_out/s18
python tools/host_profile.py 101 > gpurun_out/s18/host101.txt 2>&1
python tools/host_profile.py 1025 > gpurun_out/s18/host1025.txt 2>&1
python -m pytest tests/test_gpu_solvers.py -x -q -m gpu > gpurun_out/s18/pytest_solvers.log 2>&1
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/s18/bench.log 2>gpurun_out/s18/bench.err
echo finished
