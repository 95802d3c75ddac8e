# Produces the round's evidence: plain bench (exit 0) first, then the ncu launch list of the same command and one
# --set full capture each of the two dominant kernels.  Outputs under gpurun_out/.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "short rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sgd_rating_epoch_kernel -s 3 -c 1 -o gpurun_out/prof_sgd_r01b -f python bench.py --steps 2 --warmup 3 --no-topn --no-e2e --no-cpu-baseline > gpurun_out/ncu_sgd.log 2>&1; echo "ncu sgd rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:topn_tc_kernel -s 1 -c 1 -o gpurun_out/prof_topn_r01b -f python bench_topn.py --users 37888 --items 262144 --k 128 --steps 1 --verify 0 --cpu-sample 0 --path 2 > gpurun_out/ncu_topn.log 2>&1; echo "ncu topn rc=$?"
ls -la gpurun_out | tail -12
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
