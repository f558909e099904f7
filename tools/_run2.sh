set -x
REGIME=sparse REPS=1 python tools/prof_bwd.py > gpurun_out/plain_sparse.log 2>&1 && REGIME=sparse REPS=1 ncu --set full --clock-control none --import-source on -k regex:'bwd_bucket|dhidden|dw_kernel' -c 3 -o gpurun_out/r02_bwd_sparse_full -f python tools/prof_bwd.py > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
