cd $GRAFT_REPO_ROOT
for b in ${1:-0 1}; do
  timeout 300 ncu --set full --clock-control none -k regex:sgd_rating_epoch_kernel -s 5 -c 1 -o gpurun_out/prof_block$b -f python tools/probe_block_shape.py 8 --block=$b > gpurun_out/ncu_block$b.log 2>&1; echo "block $b rc=$?"
done
