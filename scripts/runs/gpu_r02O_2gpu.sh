# scripts/gpu_r02O_2gpu.sh — HEAD on two GPUs (the elastic kernel's new CTA shape under the boundary-first schedule):
# multi-rank tests on NVLink peers and the two-block line with its parity block
set -x
T=r02O
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_host_cpp.py tests/test_gpu_contact.py -m gpu -q -k "decomposed or decomposes or across_partitions" ) > gpurun_out/${T}_pytest_2gpu.log 2>&1; tail -5 gpurun_out/${T}_pytest_2gpu.log
bash scripts/bench_config4.sh $T 2 400
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02O_bench_config4_2gpu_twoblock.json").read().splitlines() if l.startswith("{")][-1])
print("twoblock N=2 value %.4g ms %.3f clocks %s parity %s" % (d["value"], d["ms_per_step"], d["clocks"], {k: d["parity"].get(k) for k in ("ok", "replicas_bit_equal", "max_rel_f")}))
PY
