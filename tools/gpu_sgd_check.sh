cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; tail -1 gpurun_out/bench_n1_full.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); t=d['topn']
print('sgd', d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_call'])
print('topn', t['value'], t['ms_per_step'], t['device_ms'], t['phase_ms'], t['roofline']['frac'], t['certificate'], t['parity'])"
