cd /root/repo; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_sharded.py --shard-chars 268435456 \
    > gpurun_out/r01m2_sharded_n2.json 2> gpurun_out/r01m2_sharded_n2.log
grep -v "gpu_sa" gpurun_out/r01m2_sharded_n2.log | tail -12
cat gpurun_out/r01m2_sharded_n2.json
