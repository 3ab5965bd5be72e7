set -x
timeout -s KILL 400 python -m pytest tests/test_gpu_tc_bwd.py tests/test_gpu_grad.py -q -s > gpurun_out/r02t_tcb_test4.log 2>&1; tail -5 gpurun_out/r02t_tcb_test4.log
timeout -s KILL 400 python profiles/scripts/tcb_bench.py > gpurun_out/r02t_tcb_bench2.json 2> gpurun_out/r02t_tcb_bench2.err; cat gpurun_out/r02t_tcb_bench2.json; tail -5 gpurun_out/r02t_tcb_bench2.err
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:coupling_tcb -s 3 -c 1 -o gpurun_out/r02t_coupling_tcb python profiles/scripts/tcb_bench.py --rows 262144 --steps 1 > gpurun_out/r02t_ncu_tcb.log 2>&1; tail -3 gpurun_out/r02t_ncu_tcb.log
