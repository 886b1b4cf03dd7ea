#!/bin/bash
# N-GPU evidence: the NCCL bit-exact test and bench.py exactly as the driver launches it (weak, strong, config 4)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r02_pytest_multi_n$N.log 2>&1; tail -3 gpurun_out/r02_pytest_multi_n$N.log
run() { tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N "$@" > gpurun_out/r02_bench_${tag}_n$N.json 2> gpurun_out/r02_bench_${tag}_n$N.err; tail -2 gpurun_out/r02_bench_${tag}_n$N.err | cut -c1-200; cut -c1-600 gpurun_out/r02_bench_${tag}_n$N.json; }
run weak --steps 4 --warmup 3
run strong --steps 6 --warmup 3 --scaling strong
run action512 --config action512 --steps 6 --warmup 3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/r02_bench_ref_n$N.json 2> gpurun_out/r02_bench_ref_n$N.err; cut -c1-400 gpurun_out/r02_bench_ref_n$N.json
