#!/bin/bash
mkdir -p gpurun_out
{
echo "=== nowait profile kernel 3 and 1"; MGFB_LIB=$PWD/mgf_b200/lib/libmgfb_nowait.so MGFB_AB="3:0,1:0" MGFB_AB_STEPS=4 timeout 300 python tools/solver_ab.py C2settled 2>&1 | grep -v "visits=0" | tail -8
} > gpurun_out/run3.log 2>&1
tail -60 gpurun_out/run3.log
