mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 4 --warmup 3 > gpurun_out/r02_bench_weak_n4.json 2> gpurun_out/r02_bench_weak_n4.err; tail -2 gpurun_out/r02_bench_weak_n4.err | cut -c1-200; cut -c1-500 gpurun_out/r02_bench_weak_n4.json
