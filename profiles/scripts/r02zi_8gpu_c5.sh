set -x
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout -s KILL 240 $RUN bench.py --gpus 8 --steps 10 --config c5 --scaling strong --no-e2e > gpurun_out/r02zi_c5_8gpu_strong.json 2> gpurun_out/r02zi_err.log; echo rc=$?
cut -c1-420 gpurun_out/r02zi_c5_8gpu_strong.json; tail -2 gpurun_out/r02zi_err.log
