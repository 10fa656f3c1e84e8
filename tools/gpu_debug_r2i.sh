#!/bin/bash
out=gpurun_out/r2m; mkdir -p $out
timeout 900 python tools/gpu_debug_r2i.py > $out/steps.log 2>&1
tail -n 60 $out/steps.log | cut -c1-330
