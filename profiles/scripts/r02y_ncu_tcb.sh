set -x
T='python profiles/scripts/bench_train_c3.py --steps 1 --warmup 3 --rows 262144'
timeout -s KILL 300 $T > gpurun_out/r02y_plain_train_c3.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02y_launches_train_c3.csv $T > gpurun_out/r02y_ncu_launches_train_c3.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:coupling_tcb -s 26 -c 2 -o gpurun_out/r02y_coupling_tcb $T > gpurun_out/r02y_ncu_full_tcb.log 2>&1
tail -3 gpurun_out/r02y_ncu_full_tcb.log
