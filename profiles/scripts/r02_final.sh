set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r02_final_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_gputest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_final_gputest.log
timeout 600 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"; cat gpurun_out/r02_final_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_final_bench_ref.json
