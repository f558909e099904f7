#!/usr/bin/env bash
# Round-end evidence on ONE B200 (run through gpurun): GPU test suite, the bench (both arms), ncu launch lists of the
# bench and of the native training trunk, ncu --set full of the round's new memory-bound kernels, one rank's shard of
# BASELINE config 5 with the _bulk emit.  Everything lands in gpurun_out/ (copied into profiles/ afterwards).
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02f_gpu_tests.log
python bench.py --steps 20 --warmup 5 > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02f_bench_reference_arm.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02f_launches_bench_n1.csv \
  python bench.py --steps 2 --warmup 3 --no-train-full --no-cpu-baseline --no-gpu-reference --no-ragged > $O/r02f_ncu_bench.log 2>&1
NCU=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 700 -c 2500 --csv \
  --log-file $O/r02f_launches_trunk_train.csv python tools/time_trunk_train.py > $O/r02f_ncu_trunk_train.log 2>&1
NCU=1 ncu --set full --clock-control none --import-source on -k "regex:ln_bwd|geglu_bwd|dqkv_pack|dgamma_reduce" -s 8 -c 6 \
  -o $O/r02f_prof_trunk_bwd -f python tools/time_trunk_train.py > $O/r02f_ncu_trunk_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:prologue_" -s 24 -c 6 \
  -o $O/r02f_prof_prologue -f python tools/time_prologue.py > $O/r02f_ncu_prologue.log 2>&1
python tools/time_prologue.py > $O/r02f_time_prologue.log 2>&1
python tools/time_trunk_train.py 2>/dev/null | grep shape > $O/r02f_time_trunk_train.jsonl
python tools/bench_configs.py --docs-per-gpu 125000 --native-trunk --out $O/r02f_configs_one_shard.json > $O/r02f_configs.log 2>&1
tail -2 $O/r02f_gpu_tests.log; tail -c 300 $O/r02f_bench_n1.err; ls -la $O | grep r02f | wc -l
