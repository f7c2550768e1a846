# scripts/gpu_r02d.sh — round 2, fourth GPU call: the pipelined host step (tests + the e2e figure of the default bench
# line, with and without the pipeline), the timing-log test.
set -x
T=r02d
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "step_host or pipelined or timing or bc_programs or time_dependent" ) > gpurun_out/${T}_pytest.log 2>&1; tail -15 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --no-cpu > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; echo rc=$?; tail -3 gpurun_out/${T}_bench_n400.err
timeout 900 python bench.py --no-cpu --no-parity --host-chunks 0 --steps 3 > gpurun_out/${T}_bench_n400_nopipe.json 2> gpurun_out/${T}_bench_n400_nopipe.err; echo rc=$?
timeout 900 python bench.py --no-cpu --no-parity --host-chunks 64 --steps 3 > gpurun_out/${T}_bench_n400_c64.json 2> gpurun_out/${T}_bench_n400_c64.err; echo rc=$?
python - <<'PY'
import json
for f in ("r02d_bench_n400", "r02d_bench_n400_nopipe", "r02d_bench_n400_c64"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        e = d["e2e"]
        print(f, "value %.4g  e2e %.4g  e2e ms %.1f  h2d %.1f GB/s d2h %.1f GB/s  chunks %s numa %s parity %s" % (d["value"], e["value"], e["ms_per_step"], e["h2d_gbs_per_rank"], e["d2h_gbs_per_rank"], e["host_chunks"], e["numa_binding"], d.get("parity", {}).get("ok")))
    except Exception as ex:
        print(f, "failed", ex)
PY
