#!/bin/bash
# round-2 evidence for profiles/: ncu --set full of the four hot operators and of every other kernel the path
# launches, the launch list of the bench command and of one un-graphed step, compute-sanitizer on smoke
mkdir -p gpurun_out
PROF_ITERS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_attn_tc|k_ffn_tc" -c 8 -f -o gpurun_out/r02_prof_ops python scripts/prof_ops.py qkv attn outproj_ln ffn > gpurun_out/r02_prof_ops.log 2>&1
tail -2 gpurun_out/r02_prof_ops.log
MLDB_BRANCHES=1 timeout 900 ncu --set full --clock-control none -k regex:"k_ln|k_feats2joints|k_cfg_sched|k_sched_step|k_rows_to_split|k_assemble_tokens|k_timestep_features|k_mem_tokens|k_permute_01|k_f32_to_split|k_split_to_f32|k_rows_out_permuted|k_gemm_simt|k_dup_lengths|k_gather_rows|k_step" -c 80 -f -o gpurun_out/r02_prof_misc python scripts/prof_step.py > gpurun_out/r02_prof_misc.log 2>&1
tail -2 gpurun_out/r02_prof_misc.log
MLDB_BRANCHES=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_prof_step.csv python scripts/prof_step.py > gpurun_out/r02_prof_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_bench_under_ncu.log 2> gpurun_out/r02_bench_under_ncu.err
wc -l gpurun_out/r02_launches_bench.csv gpurun_out/r02_launches_prof_step.csv
timeout 1200 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck.log
