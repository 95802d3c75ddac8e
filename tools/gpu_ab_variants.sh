# A/B of library builds: tools/gpu_ab_variants.sh "b a c" ; each librec_b200/_lib/variants/<v>.so is copied over the
# in-tree library, then C2 (bench.py --no-topn), C3 and C4 (bench_configs.py) are timed.
cd $GRAFT_REPO_ROOT
cp librec_b200/_lib/liblibrec_b200.so /tmp/main.so
for v in ${1:-b}; do
  cp librec_b200/_lib/variants/$v.so librec_b200/_lib/liblibrec_b200.so
  if [ -z "$SKIP_C2" ]; then python bench.py --steps 10 --warmup 3 --no-topn --no-e2e --no-cpu-baseline > gpurun_out/ab_${v}_c2.json 2> gpurun_out/ab_${v}_c2.err; fi
  python bench_configs.py ${2:-} > gpurun_out/ab_${v}_cfg.json 2> gpurun_out/ab_${v}_cfg.err
  python - <<PY
import json
import os
if not os.environ.get("SKIP_C2"):
  d=json.loads(open("gpurun_out/ab_${v}_c2.json").read().strip().splitlines()[-1])
  print("$v C2 %.3f G/s  step %.3f ms kernel %.3f ms loss %.0f" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["final_loss"]))
for l in open("gpurun_out/ab_${v}_cfg.json"):
    c=json.loads(l); print("$v", c["config"], "%.3f G/s  %.3f ms" % (c["value"]/1e9, c["ms_per_step"]), "loss %.0f" % c["losses"][-1], c["safeguard"])
PY
done
cp /tmp/main.so librec_b200/_lib/liblibrec_b200.so
