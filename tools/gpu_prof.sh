#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_kernels python tools/prof_kernels.py > gpurun_out/prof_kernels.log 2>&1
tail -3 gpurun_out/prof_kernels.log
ls -la gpurun_out/*.ncu-rep
