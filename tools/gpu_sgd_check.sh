# SGD regression pass on one GPU: all GPU tests, C2 bench (SGD leg), C3, C4 twice (stability at the edge), block probe
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-topn > gpurun_out/bench_n1_sgd.json 2> gpurun_out/bench_n1_sgd.err; echo "bench rc=$?"
for i in 1 2; do LRK_SGD_TRACE=1 python bench_configs.py ${1:---only c4} > gpurun_out/configs_run$i.json 2> gpurun_out/configs_run$i.err; done
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1_sgd.json").read().strip().splitlines()[-1])
print("C2 %.3f G/s step %.3f ms kernel %.3f ms loss %.0f e2e %.3f G/s (%.2f ms/call)" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["final_loss"], d["e2e"]["value"]/1e9, d["e2e"]["ms_per_call"]))
for i in (1,2):
    for l in open("gpurun_out/configs_run%d.json"%i):
        c=json.loads(l); print(i, c["config"], "%.3f G/s  %.3f ms" % (c["value"]/1e9, c["ms_per_step"]), "loss %.0f" % c["losses"][-1], c["safeguard"])
PY
grep "^\[sgd\]" gpurun_out/configs_run1.err | head -8
bash tools/gpu_probe_flush.sh 512 "1 8" 
