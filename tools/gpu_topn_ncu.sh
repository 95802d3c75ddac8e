set -x
cd $GRAFT_REPO_ROOT
B="python bench_topn.py --users 37888 --items 262144 --k 128 --steps 1 --verify 0 --cpu-sample 0 --path 2"
timeout 300 $B > gpurun_out/t_small.json 2>&1; tail -1 gpurun_out/t_small.json
timeout 900 ncu --set full --import-source on --clock-control none -k regex:topn_tc_kernel -s 1 -c 1 -o gpurun_out/prof_topn6 -f $B > gpurun_out/ncu_topn6.log 2>&1
tail -3 gpurun_out/ncu_topn6.log
