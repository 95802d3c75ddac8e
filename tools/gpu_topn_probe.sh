set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -1
timeout 600 python -m pytest tests/test_gpu_topn.py -x -q 2>&1 | tail -5
timeout 600 python bench_topn.py --users 1048576 --items 1048576 --k 128 --steps 2 --verify 64 --cpu-sample 64 > gpurun_out/topn_c5.json 2> gpurun_out/topn_c5.err; tail -1 gpurun_out/topn_c5.json
timeout 600 python bench_topn.py --users 138493 --items 26744 --k 128 --steps 3 --verify 256 --cpu-sample 256 > gpurun_out/topn_c3.json 2> gpurun_out/topn_c3.err; tail -1 gpurun_out/topn_c3.json
