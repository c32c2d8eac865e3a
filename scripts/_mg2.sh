export ANQS_ALLOC_TRACE=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/bench_vmc_sharded.py --steps 12 2>&1 | grep "iteration \|per-iteration" | cut -c1-200
echo "---- expandable segments"
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/bench_vmc_sharded.py --steps 12 2>&1 | grep "iteration \|per-iteration" | cut -c1-200
