set -x
nvidia-smi --query-gpu=name --format=csv
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -x -q -s > gpurun_out/r02f_dist_test.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02f_dist_test.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TNF_PEER_EXCHANGE=0 timeout 300 $RUN bench.py --gpus 2 --steps 10 --no-fp32 --scaling strong > gpurun_out/r02f_bench_2gpu_strong_nccl.json 2> gpurun_out/r02f_err1.log; echo rc=$?
TNF_PEER_EXCHANGE=1 timeout 300 $RUN bench.py --gpus 2 --steps 10 --no-fp32 --scaling strong > gpurun_out/r02f_bench_2gpu_strong_peer.json 2> gpurun_out/r02f_err2.log; echo rc=$?
TNF_PEER_EXCHANGE=1 timeout 300 $RUN bench.py --gpus 2 --steps 10 --no-fp32 > gpurun_out/r02f_bench_2gpu_weak_peer.json 2> gpurun_out/r02f_err3.log; echo rc=$?
timeout 300 $RUN profiles/scripts/bench_train.py --steps 5 > gpurun_out/r02f_train_2gpu.json 2> gpurun_out/r02f_err4.log; echo rc=$?
for f in gpurun_out/r02f_bench_2gpu_strong_nccl.json gpurun_out/r02f_bench_2gpu_strong_peer.json gpurun_out/r02f_bench_2gpu_weak_peer.json gpurun_out/r02f_train_2gpu.json; do echo $f; cut -c1-400 $f; done
tail -3 gpurun_out/r02f_err2.log
