#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/chunk.log
for c in 0 256 192 128 96 64; do
  echo "== MLDB_CHUNK=$c" >> gpurun_out/chunk.log
  MLDB_CHUNK=$c timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "device-resident|layer" >> gpurun_out/chunk.log
done
cat gpurun_out/chunk.log
