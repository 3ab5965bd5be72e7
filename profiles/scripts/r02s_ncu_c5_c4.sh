set -x
C5="python bench.py --config c5 --batch 262144 --steps 1 --warmup 3 --no-e2e"
$C5 > gpurun_out/r02s_plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02s_launches_c5.csv $C5 > gpurun_out/r02s_ncu_launches_c5.log 2>&1
$C5 > gpurun_out/r02s_plain_c5b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coupling_tc6 -s 54 -c 2 -o gpurun_out/r02s_coupling_tc6_c5 $C5 > gpurun_out/r02s_ncu_full_c5.log 2>&1
C4="python profiles/scripts/bench_train.py --steps 1 --warmup 3"
$C4 > gpurun_out/r02s_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:coupling_generic -s 12 -c 4 -o gpurun_out/r02s_coupling_generic_c4 $C4 > gpurun_out/r02s_ncu_full_c4.log 2>&1
ls -la gpurun_out | grep r02s
