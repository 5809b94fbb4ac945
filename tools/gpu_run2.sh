#!/bin/bash
mkdir -p gpurun_out
{
echo "=== quick parity (local solver default)"; timeout 600 python -m pytest tests/test_gpu_step.py -x -q -m gpu -k "c1 or jittered or snapshot or pipelined or handover or capsules" 2>&1 | tail -15
echo "=== A/B"; timeout 600 python tools/solver_ab.py C2pile C2settled 2>&1 | tail -20
echo "=== per-visit profile, C2settled"; MGFB_LIB=$PWD/mgf_b200/lib/libmgfb_prof.so MGFB_AB="1:0,3:0" MGFB_AB_STEPS=4 timeout 300 python tools/solver_ab.py C2settled 2>&1 | grep -v "visits=0" | tail -20
} > gpurun_out/run2.log 2>&1
tail -60 gpurun_out/run2.log
