set -x
nvidia-smi --query-gpu=name --format=csv | head -3
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 240 $RUN bench.py --gpus 8 --steps 10 --no-fp32 --no-e2e --scaling strong > gpurun_out/r02g_c3_8gpu_strong_peer.json 2> gpurun_out/r02g_err1.log; echo rc=$?
TNF_PEER_EXCHANGE=0 timeout 240 $RUN bench.py --gpus 8 --steps 10 --no-fp32 --no-e2e --scaling strong > gpurun_out/r02g_c3_8gpu_strong_nccl.json 2> gpurun_out/r02g_err2.log; echo rc=$?
timeout 240 $RUN bench.py --gpus 8 --steps 10 --no-fp32 > gpurun_out/r02g_c3_8gpu_weak_peer.json 2> gpurun_out/r02g_err3.log; echo rc=$?
timeout 240 $RUN bench.py --gpus 8 --steps 10 --config c5 --scaling strong --no-e2e > gpurun_out/r02g_c5_8gpu_strong.json 2> gpurun_out/r02g_err4.log; echo rc=$?
timeout 240 $RUN profiles/scripts/bench_train.py --steps 5 > gpurun_out/r02g_train_8gpu.json 2> gpurun_out/r02g_err5.log; echo rc=$?
timeout 200 $RUN profiles/microbench/pcie_concurrent.py > gpurun_out/r02g_pcie_8.txt 2> gpurun_out/r02g_err6.log; echo rc=$?
for f in gpurun_out/r02g_*.json gpurun_out/r02g_pcie_8.txt; do echo $f; grep -v "^NCCL" $f | cut -c1-330; done
tail -3 gpurun_out/r02g_err4.log gpurun_out/r02g_err6.log
