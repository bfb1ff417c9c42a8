python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|Error|rgb_over|assert" | cut -c1-330
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "golden or strips or immediate or empty" 2>&1 | tail -6
