set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02c_gputest.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
cat gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
