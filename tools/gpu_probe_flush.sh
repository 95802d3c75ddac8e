cd $GRAFT_REPO_ROOT
for hf in ${1:-0 16384}; do
  LRK_SGD_HOT_FLUSH=$hf timeout 200 python tools/probe_block_shape.py ${2:-8} > gpurun_out/probe_hf$hf.json 2> gpurun_out/probe_hf$hf.err
  echo "HOT_FLUSH=$hf rc=$?"; tail -2 gpurun_out/probe_hf$hf.err | cut -c1-300
  python - <<PY
import json
for l in open("gpurun_out/probe_hf$hf.json"):
    d=json.loads(l); print(d["G"], d["partition"], [round(x,3) for x in d["kernel_ms"]], "sum", round(d["sum_ms"],3), "Gxmax", round(d["G_x_max_ms"],3), "rollbacks", d["rollbacks"], "loss", [round(x) for x in d["loss_8"]])
PY
done
