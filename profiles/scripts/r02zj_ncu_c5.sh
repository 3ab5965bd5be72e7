set -x
C5="python bench.py --config c5 --batch 262144 --steps 1 --warmup 3 --no-e2e"
timeout -s KILL 200 $C5 > gpurun_out/r02zj_plain_c5.log 2>&1 && \
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:coupling_tc6 -s 54 -c 2 -o gpurun_out/r02zj_coupling_tc6_c5 $C5 > gpurun_out/r02zj_ncu_full_c5.log 2>&1
tail -2 gpurun_out/r02zj_ncu_full_c5.log
