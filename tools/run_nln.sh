#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host_layer.py -m gpu -x -q -k "nln or cartpole_example" 2>&1 | tail -25
