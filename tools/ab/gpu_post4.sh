set -x
timeout 300 python -m pytest tests/test_prepost_gpu.py -x -q -m gpu -k "contour" 2>&1 | tail -6
timeout 200 python tools/bench_prepost.py 2>&1 | grep "contour"
