cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_topn.py -x -q 2>&1 | tail -3
B="python bench_topn.py --users 37888 --items 262144 --k 128 --steps 3 --verify 64 --cpu-sample 0 --path 2"
timeout 300 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['phase_ms'], d['roofline']['frac'], d['parity'], d['certificate']['fallback_users'], d['certificate']['resweep_users'])"
timeout 600 python bench_topn.py --users 1048576 --items 1048576 --k 128 --steps 2 --verify 64 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phase_ms'], d['roofline']['frac'], d.get('parity'), d['certificate']['fallback_users'], d['certificate']['resweep_users'])"
timeout 600 python bench_topn.py --users 138493 --items 26744 --k 128 --steps 3 --verify 256 --cpu-sample 0 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['phase_ms'], d['roofline']['frac'], d.get('parity'), d['certificate']['fallback_users'], d['certificate']['resweep_users'])"
