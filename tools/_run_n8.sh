set -x
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo bench8 rc=$?
tail -c 600 gpurun_out/r02_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/bench_configs.py --docs-per-gpu 125000 --native-trunk --out gpurun_out/r02_configs_3_4b_5_n8_1M_docs.json > gpurun_out/r02_configs_n8.log 2>&1; echo cfg rc=$?
tail -c 400 gpurun_out/r02_configs_n8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo bench4 rc=$?
