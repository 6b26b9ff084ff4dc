timeout 900 python -m pytest tests/test_gpu_crop.py tests/test_gpu_cascade.py -x -q -m gpu 2>&1 | tail -2
HGSFA_DETECT_PROFILE=1 timeout 600 python tools/bench_detect.py --cpu-images 0 2>&1 | tail -1 | cut -c1-200
