python -m pytest tests/test_gpu_parity.py tests/test_cpp_host_layer.py -q -m gpu -k "tsallis or racer_colored" 2>&1 | tail -8
