#!/bin/bash
# GPU-box script: quadrotor parity tests, C++ example, smoke, and a short bench line for the C = 4 pair.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host_layer.py -m gpu -x -q -k "quadrotor" 2>&1 | tail -15
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8
timeout 300 python bench.py --workload quadrotor --rollouts 32768 --timesteps 100 --steps 200 --warmup 20 2>&1 | tail -2 | tee gpurun_out/bench_quadrotor.json
