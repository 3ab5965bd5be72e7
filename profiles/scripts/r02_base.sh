set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest0.log 2>&1; echo "pytest rc=$?"
timeout 400 python bench.py --steps 10 > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --precision fp32 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench0_fp32.json 2> gpurun_out/r02_bench0_fp32.err; echo "bench fp32 rc=$?"
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/r02_gputest0.log; cat gpurun_out/r02_bench0.json; cat gpurun_out/r02_bench0_fp32.json; tail -5 gpurun_out/r02_memcheck_smoke.log
