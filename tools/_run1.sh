set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/time_bwd.py 2>&1 | grep -E "^(dense|sparse)" | cut -c1-520 | tee gpurun_out/r02_time_bwd.log
SPLADE_B200_BWD_CONCURRENT=0 python tools/time_bwd.py 2>&1 | grep -E "^(dense|sparse)" | cut -c1-520 | tee gpurun_out/r02_time_bwd_serial.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train-full --no-full-model --no-ragged --no-gpu-reference > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; echo bench rc=$?
