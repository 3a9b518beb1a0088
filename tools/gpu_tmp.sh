timeout 600 python -m pytest tests/test_gpu_cycle_multi.py -x -q -m gpu -k legacy_segments 2>&1 | grep -E "^E|assert|Error" | head -20
