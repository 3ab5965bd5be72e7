set -x
timeout 300 python -m pytest tests/test_gpu_cde_fused.py -x -q > gpurun_out/r02d_cde_test.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02d_cde_test.log
timeout 300 python profiles/scripts/bench_configs.py c2b c4like > gpurun_out/r02d_configs.jsonl 2> gpurun_out/r02d_configs.err; echo "rc=$?"
cat gpurun_out/r02d_configs.jsonl; tail -5 gpurun_out/r02d_configs.err
