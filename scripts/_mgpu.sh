N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_${N}gpu.json | head -c 300; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/bench_vmc_sharded.py --steps 10 > gpurun_out/r2_vmc_c5_${N}gpu.json 2> gpurun_out/r2_vmc_c5_${N}gpu.err; echo "vmc rc=$?"
cut -c1-200 gpurun_out/r2_vmc_c5_${N}gpu.json
