N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_vmc_sharded.py --steps 14 > gpurun_out/r2_vmc_c5_${N}gpu.json 2> gpurun_out/r2_vmc_c5_${N}gpu.err; echo "vmc rc=$?"
grep "per-iteration" gpurun_out/r2_vmc_c5_${N}gpu.err | cut -c1-200
grep "^{" gpurun_out/r2_vmc_c5_${N}gpu.json | cut -c1-160
