cd $GRAFT_REPO_ROOT
LRK_DSGD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 5 --warmup 3 --no-topn > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
grep "dsgd rank 0" gpurun_out/bench_n4.err | tail -2
tail -1 gpurun_out/bench_n4.json | cut -c1-300
