set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gputest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02e_gputest.log
timeout 300 python profiles/scripts/bench_configs.py c5 > gpurun_out/r02e_c5.jsonl 2> gpurun_out/r02e_c5.err; cat gpurun_out/r02e_c5.jsonl; tail -3 gpurun_out/r02e_c5.err
timeout 600 python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
cat gpurun_out/r02e_bench.json; tail -3 gpurun_out/r02e_bench.err
