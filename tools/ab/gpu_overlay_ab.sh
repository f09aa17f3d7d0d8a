set -x
timeout 300 python -m pytest tests/test_prepost_gpu.py -x -q -m gpu -k "overlay or gray or fold" 2>&1 | tail -8
OCTSEG_OVERLAY_IMPL=ty32 timeout 300 python -m pytest tests/test_prepost_gpu.py -x -q -m gpu -k "overlay" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_pipeline_gpu.py -x -q -m gpu -k "fold or save_results" 2>&1 | tail -8
timeout 100 python tools/bench_prepost.py 2>&1 | tail -2
OCTSEG_OVERLAY_IMPL=ty32 timeout 100 python tools/bench_prepost.py 2>&1 | tail -1
