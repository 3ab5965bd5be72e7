set -x
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload train_c3 --steps 5 --warmup 3 > gpurun_out/r02z_train_c3_8gpu.json 2> gpurun_out/r02z_train_c3_8gpu.err
cat gpurun_out/r02z_train_c3_8gpu.json; tail -3 gpurun_out/r02z_train_c3_8gpu.err
