set -x
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-fp32"
$CMD > gpurun_out/r02b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_launches.log 2>&1
$CMD > gpurun_out/r02b_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coupling_tc5 -s 22 -c 4 -o gpurun_out/r02b_coupling_tc5 $CMD > gpurun_out/r02b_ncu_full.log 2>&1
CMD6="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-fp32 --precision fp32"
$CMD6 > gpurun_out/r02b_plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coupling_tc6 -s 23 -c 2 -o gpurun_out/r02b_coupling_tc6 $CMD6 > gpurun_out/r02b_ncu_full6.log 2>&1
CMDC="python profiles/scripts/cde_kernel_bench.py"
$CMDC > gpurun_out/r02b_plain_cde.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cde_logprob_tc -s 3 -c 1 -o gpurun_out/r02b_cde_logprob_tc $CMDC > gpurun_out/r02b_ncu_full_cde.log 2>&1
ls -la gpurun_out/ | grep r02b_
