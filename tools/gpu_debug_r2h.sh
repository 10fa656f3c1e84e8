#!/bin/bash
out=gpurun_out/r2h; mkdir -p $out
timeout 600 python tools/gpu_debug_r2h.py > $out/canon.log 2>&1
QB200_NO_CHOLQR=1 timeout 600 python tools/gpu_debug_r2h.py > $out/canon_nochol.log 2>&1
for f in canon canon_nochol; do echo "== $f"; tail -n 12 $out/$f.log | cut -c1-300; done
