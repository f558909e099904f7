set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-train --no-full-model > gpurun_out/r02_bench_n2_tf.json 2> gpurun_out/r02_bench_n2_tf.err; echo bench rc=$?
SPLADE_B200_DEFER_DW=0 SPLADE_B200_LAZY_TOKEN_WEIGHTS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-train --no-full-model > gpurun_out/r02_bench_n2_tf_nodefer.json 2> gpurun_out/r02_bench_n2_tf_nodefer.err; echo bench rc=$?
