#!/bin/bash
# Round 2: interactive frames in one launch (xf_substep_varying), Sim layer, full GPU suite.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/gpu_tests_sim.log
cat gpurun_out/gpu_tests_sim.log
