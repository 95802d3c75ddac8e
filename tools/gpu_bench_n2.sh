cd $GRAFT_REPO_ROOT
N=${1:-2}
LRK_DSGD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-topn > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
grep "dsgd rank 0" gpurun_out/bench_n$N.err | tail -1 | cut -c1-900
grep -i "error\|exception" gpurun_out/bench_n$N.err | head -3
tail -1 gpurun_out/bench_n$N.json | cut -c1-330
