set -x
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:coupling_tcb -s 3 -c 1 -o gpurun_out/r02u_coupling_tcb python profiles/scripts/tcb_bench.py --rows 262144 --steps 1 > gpurun_out/r02u_ncu_tcb.log 2>&1; tail -3 gpurun_out/r02u_ncu_tcb.log
