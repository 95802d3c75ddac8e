cd $GRAFT_REPO_ROOT
B="python bench_topn.py --users 37888 --items 262144 --k 128 --steps 1 --verify 0 --cpu-sample 0 --path 2"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:topn_tc_kernel -s 1 -c 1 -o gpurun_out/prof_topn8 -f $B > gpurun_out/ncu_topn8.log 2>&1
tail -2 gpurun_out/ncu_topn8.log
