set -x
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-fp32"
$CMD > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
$CMD > gpurun_out/r02_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coupling_tc5 -s 24 -c 2 -o gpurun_out/r02_coupling_tc5 $CMD > gpurun_out/r02_ncu_full.log 2>&1
CMD6="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-fp32 --precision fp32"
$CMD6 > gpurun_out/r02_plain6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coupling_tc6 -s 24 -c 2 -o gpurun_out/r02_coupling_tc6 $CMD6 > gpurun_out/r02_ncu_full6.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log gpurun_out/r02_ncu_full6.log
ls -la gpurun_out/
