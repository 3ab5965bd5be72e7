set -x
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train_c3 --steps 5 --warmup 3 > gpurun_out/r02z_train_c3_2gpu.json 2> gpurun_out/r02z_train_c3_2gpu.err
cat gpurun_out/r02z_train_c3_2gpu.json; tail -3 gpurun_out/r02z_train_c3_2gpu.err
