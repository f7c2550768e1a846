set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/G_smoke.log 2>&1; tail -2 gpurun_out/G_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/G_pytest.log 2>&1; tail -4 gpurun_out/G_pytest.log
timeout 900 python bench.py > gpurun_out/G_bench_n400.json 2> gpurun_out/G_bench_n400.err; tail -3 gpurun_out/G_bench_n400.err
python - <<PY
import json
d=json.loads(open("gpurun_out/G_bench_n400.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f e2e %.4g launches %d"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["gpu_launches"]))
PY
