# scripts/gpu_r02h.sh — round 2, one GPU: the state kernel with coalesced record staging (tests + throughput), the
# ORDERED bench line of the headline configuration, the host<->device copy ceiling of the box.
set -x
T=r02h
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "state or pipelined or binding or step_host" ) > gpurun_out/${T}_pytest.log 2>&1; tail -8 gpurun_out/${T}_pytest.log
for N in 200 400; do
  timeout 900 python bench.py --n $N --material j2_plasticity --steps 10 --no-cpu --no-e2e > gpurun_out/${T}_bench_j2_n$N.json 2> gpurun_out/${T}_bench_j2_n$N.err; echo rc=$?
done
timeout 900 python bench.py --assembly ordered --no-cpu --copy-ceiling > gpurun_out/${T}_bench_n400_ordered.json 2> gpurun_out/${T}_bench_n400_ordered.err; echo rc=$?; tail -2 gpurun_out/${T}_bench_n400_ordered.err
python - <<'PY'
import json
for f in ("r02h_bench_j2_n200", "r02h_bench_j2_n400", "r02h_bench_n400_ordered"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms %.3f kernel %.3f fp64 frac %.4f  node ms %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["fp64"]["frac"], d["node_kernels_ms"], d["clocks"]))
        if d.get("e2e"): print("   e2e", d["e2e"]["value"], d["e2e"].get("copy_ceiling"))
        print("   parity", (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("max_rel_f"))
    except Exception as ex:
        print(f, "failed", ex)
PY
