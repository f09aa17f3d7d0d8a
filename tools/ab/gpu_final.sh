set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gpu_tests_s3.log; cat gpurun_out/gpu_tests_s3.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; cat gpurun_out/bench_s3.json
