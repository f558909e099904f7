set -x
python -m pytest tests/test_gpu_loss.py tests/test_gpu_peer.py -x -q 2>&1 | tail -5
python tools/time_kernels.py 2>&1 | grep -i "loss" | tee gpurun_out/r02_time_loss.log
