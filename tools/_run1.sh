set -x
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo bench rc=$?
tail -c 300 gpurun_out/r02_bench_n1_c.err
python -m pytest tests/test_gpu_model.py tests/test_gpu_loss.py -x -q 2>&1 | tail -3
timeout 300 python tools/prof_peer.py > gpurun_out/plain_peer.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'peer_' -o gpurun_out/r02_peer_full -f python tools/prof_peer.py > gpurun_out/ncu_peer.log 2>&1; echo peer rc=$?
tail -3 gpurun_out/plain_peer.log; tail -3 gpurun_out/ncu_peer.log
