set -x
timeout 300 python -m pytest tests/test_prepost_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu -k "overlay or save_results" 2>&1 | tail -4
timeout 200 python tools/bench_prepost.py 2>&1 | grep "overlay"
