#!/bin/bash
# multi-GPU check (run with gpurun --gpus 2): both bench arms under torchrun, short
out=gpurun_out/${1:-r2m}; mkdir -p $out
N=${2:-2}
nvidia-smi --query-gpu=name --format=csv,noheader > $out/gpus.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 1 > $out/bench_n$N.log 2> $out/bench_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $out/ref_n$N.log 2> $out/ref_n$N.err
tail -c 1500 $out/bench_n$N.err; head -c 400 $out/bench_n$N.log; echo; head -c 300 $out/ref_n$N.log
