set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gpu_tests_s4.log; cat gpurun_out/gpu_tests_s4.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s4.json 2> gpurun_out/bench_s4.err; cut -c1-330 gpurun_out/bench_s4.json
