export ANQS_ALLOC_TRACE=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_vmc_sharded.py --steps 14 > gpurun_out/r2_vmc_c5_4gpu.json 2> gpurun_out/r2_vmc_c5_4gpu.err
grep "rank 3 iteration\|per-iteration" gpurun_out/r2_vmc_c5_4gpu.err | cut -c1-200
cut -c1-160 gpurun_out/r2_vmc_c5_4gpu.json
