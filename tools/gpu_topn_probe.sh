cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_topn.py -x -q 2>&1 | tail -5
for keep in 24 16 12; do
echo "=== keep=$keep"
export LRK_TC_KEEP=$keep
timeout 300 python bench_topn.py --users 37888 --items 262144 --k 128 --steps 2 --verify 64 --cpu-sample 0 --path 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['phase_ms'], d['roofline']['frac'], d.get('parity'), d['certificate']['fallback_users'])"
timeout 600 python bench_topn.py --users 1048576 --items 1048576 --k 128 --steps 2 --verify 64 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phase_ms'], d['roofline']['frac'], d.get('parity'), d['certificate']['fallback_users'])"
timeout 600 python bench_topn.py --users 138493 --items 26744 --k 128 --steps 3 --verify 256 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phase_ms'], d['roofline']['frac'], d.get('parity'), d['certificate']['fallback_users'])"
done
