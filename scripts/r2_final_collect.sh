#!/bin/bash
# Final round-2 evidence on one B200 (run under gpurun): GPU tests, timings of the network / sampler kernels after the move
# to the FP64 tensor-core path, the bench line, the ncu launch list of the bench, full ncu captures of the dominant kernels.
python -m pytest tests -m gpu -q 2>&1 | tail -2
python scripts/vmc_c5_phases.py 1048576 MADE 2>&1 | grep "rows\|anqs::" | cut -c1-250 > gpurun_out/r2_made_phases.txt
python scripts/sampler_c5_phases.py 2>&1 | grep "samples ->\|anqs::" | cut -c1-250 > gpurun_out/r2_sampler_phases.txt
python scripts/bench_c3_transformer.py > gpurun_out/r2_c3_transformer.json 2> gpurun_out/c3.err
python scripts/batch_reduce_time.py 1048576 > gpurun_out/r2_batch_reduce.txt 2>&1
./scripts/microbench_dmma > gpurun_out/r2_dmma_rate.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_eloc_bs -c 1 -o gpurun_out/prof_r2_fused -f python scripts/fused_prof.py 1048576 > gpurun_out/ncu_r2_fused.log 2>&1; tail -1 gpurun_out/ncu_r2_fused.log
ncu --set full --clock-control none --import-source on -k regex:transformer_backward_kernel -c 2 -o gpurun_out/prof_r2_tfm -f python scripts/tfm_bwd_prof.py 10000 0 > gpurun_out/ncu_r2_tfm.log 2>&1; tail -1 gpurun_out/ncu_r2_tfm.log
ncu --set full --clock-control none --import-source on -k regex:"made_forward_kernel|made_backward_kernel|batch_reduce_gemm|made_phase_output_kernel" -c 4 -o gpurun_out/prof_r2_made -f python scripts/vmc_c5_phases.py 262144 MADE > gpurun_out/ncu_r2_made.log 2>&1; tail -1 gpurun_out/ncu_r2_made.log
