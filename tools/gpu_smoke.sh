#!/bin/bash
# last sanity of the in-tree library: smoke() + a few fast tests across the subsystems
OUT=gpurun_out/${1:-smoke}; mkdir -p $OUT
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py tests/test_gpu_gradients.py -m gpu -q -x -k "pencil_forces or float64 or fused_adjoint or pipelined or golden" 2>&1 | tail -3
