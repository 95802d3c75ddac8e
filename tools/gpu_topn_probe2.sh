cd $GRAFT_REPO_ROOT
B="python bench_topn.py --users 37888 --items 262144 --k 128 --steps 2 --verify 0 --cpu-sample 0"
for mode in 5 6 2 1; do
  echo "== debug=$mode"; LRK_TC_DEBUG=$mode timeout 300 $B --path 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['phase_ms']['sweep'], d['roofline']['frac'])"
done
