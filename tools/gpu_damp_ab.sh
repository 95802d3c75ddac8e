cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -q -m gpu -k "sgd or host_job or dsgd" 2>&1 | tail -3
for nd in 1 0; do
echo "== NODAMP=$nd N=1"
LRK_SGD_NODAMP=$nd timeout 600 python bench.py --steps 10 --warmup 3 --no-topn --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['final_loss'])"
echo "== NODAMP=$nd N=4"
LRK_SGD_NODAMP=$nd LRK_DSGD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --no-topn > gpurun_out/damp$nd.json 2> gpurun_out/damp$nd.err
tail -1 gpurun_out/damp$nd.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['final_loss'])"
grep "dsgd rank 0" gpurun_out/damp$nd.err | tail -1 | cut -c1-420
done
