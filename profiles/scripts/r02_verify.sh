set -x
date
timeout -s KILL 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02v_gputests.log 2>&1; tail -5 gpurun_out/r02v_gputests.log
date
timeout -s KILL 600 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; tail -c 1500 gpurun_out/r02v_bench.json; tail -3 gpurun_out/r02v_bench.err
date
