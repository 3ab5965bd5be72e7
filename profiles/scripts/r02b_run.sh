set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02b_gputest.log
timeout 600 python bench.py --steps 10 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
cat gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
