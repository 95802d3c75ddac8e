cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_topn.py -x -q 2>&1 | tail -5
B="python bench_topn.py --users 37888 --items 262144 --k 128 --steps 2 --verify 64 --cpu-sample 0"
for mode in 2 1 0; do
  echo "== debug=$mode"; LRK_TC_DEBUG=$mode timeout 300 $B --path 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['phase_ms']['sweep'], d['roofline']['frac'], d.get('parity'), d['certificate'])"
done
timeout 600 python bench_topn.py --users 1048576 --items 1048576 --k 128 --steps 2 --verify 64 --cpu-sample 64 > gpurun_out/topn_c5.json 2> gpurun_out/topn_c5.err; tail -1 gpurun_out/topn_c5.json
timeout 600 python bench_topn.py --users 138493 --items 26744 --k 128 --steps 3 --verify 256 --cpu-sample 256 > gpurun_out/topn_c3.json 2> gpurun_out/topn_c3.err; tail -1 gpurun_out/topn_c3.json
