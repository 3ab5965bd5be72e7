set -x
timeout -s KILL 400 python -m pytest tests/test_gpu_tc_bwd.py -q -s > gpurun_out/r02zh_tcb_test.log 2>&1; tail -4 gpurun_out/r02zh_tcb_test.log
timeout -s KILL 400 python profiles/scripts/tcb_bench.py > gpurun_out/r02zh_tcb_bench.json 2> gpurun_out/r02zh_tcb_bench.err; cat gpurun_out/r02zh_tcb_bench.json; tail -5 gpurun_out/r02zh_tcb_bench.err
timeout -s KILL 400 python bench.py --workload train_c3 --steps 5 --warmup 3 > gpurun_out/r02zh_train_c3.json 2> gpurun_out/r02zh_train_c3.err; cat gpurun_out/r02zh_train_c3.json; tail -5 gpurun_out/r02zh_train_c3.err
