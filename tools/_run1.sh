set -x
python -m pytest tests/test_gpu_head.py tests/test_gpu_errors.py -x -q 2>&1 | tail -15
python tools/time_bwd.py 2>&1 | tee gpurun_out/r02_time_bwd.log | tail -3
for R in sparse dense; do
REGIME=$R python tools/prof_bwd.py > gpurun_out/plain_$R.log 2>&1 && REGIME=$R ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:'bwd_bucket|dhidden|dw_kernel' --csv --log-file gpurun_out/r02_bwd_launches_$R.csv python tools/prof_bwd.py > gpurun_out/ncu_$R.log 2>&1
done
