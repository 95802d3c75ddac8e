cd $GRAFT_REPO_ROOT
N=${1:-8}
EXTRA=${2:-}
LRK_DSGD_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
grep "dsgd rank 0" gpurun_out/bench_n$N.err | tail -1 | cut -c1-1500
grep -i "error\|exception" gpurun_out/bench_n$N.err | head -3
tail -1 gpurun_out/bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); t=d.get('topn')
print('sgd', d['value'], d['ms_per_step'], d['roofline']['frac'], d['config']['final_loss'])
if t: print('topn', t['value'], t['ms_per_step'], t['device_ms'], t['phase_ms'], t['roofline']['frac'], t['certificate'], t['parity'])"
