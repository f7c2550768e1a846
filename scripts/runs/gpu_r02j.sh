# scripts/gpu_r02j.sh — round 2, one GPU: host-step chunk sweep, the FMA-contracted build (deviation table + speed;
# never the headline), stamped ncu traffic of the final kernels, the closing default bench line and GPU suite.
set -x
T=r02j
mkdir -p gpurun_out
for CH in 4 8 32; do
  timeout 600 python bench.py --no-cpu --no-parity --host-chunks $CH --steps 3 > gpurun_out/${T}_bench_n400_c$CH.json 2> gpurun_out/${T}_bench_n400_c$CH.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench_n400_c$CH.json").read().strip().splitlines()[-1]); e = d["e2e"]
print("CHUNKS $CH  e2e %.4g  %.1f ms  h2d %.1f d2h %.1f GB/s" % (e["value"], e["ms_per_step"], e["h2d_gbs_per_rank"], e["d2h_gbs_per_rank"]))
PY
done
python scripts/fma_deviation.py > gpurun_out/${T}_deviation_parity_build.json 2> gpurun_out/${T}_deviation_parity_build.err
NSM_B200_LIB=nimblesm_b200/lib/variants/libnsm_b200_fma.so python scripts/fma_deviation.py > gpurun_out/${T}_deviation_fma_build.json 2> gpurun_out/${T}_deviation_fma_build.err
for MAT in neohookean elastic; do
  NSM_B200_LIB=nimblesm_b200/lib/variants/libnsm_b200_fma.so timeout 300 python bench.py --n 200 --material $MAT --steps 20 --no-e2e --no-cpu \
    > gpurun_out/${T}_bench_fma_n200_$MAT.json 2> gpurun_out/${T}_bench_fma_n200_$MAT.err
done
python - <<'PY'
import json
for f in ("r02j_deviation_parity_build", "r02j_deviation_fma_build"):
    d = json.load(open("gpurun_out/%s.json" % f)); print(f, d["dp_per_element"])
    for r in d["rows"]: print("   %-10s eps %.0e  force %.2e  sigma %.2e  F %.2e" % (r["material"], r["eps"], r["force_rel"], r["sigma_rel"], r["F_rel"]))
for m in ("neohookean", "elastic"):
    try:
        d = json.loads(open("gpurun_out/r02j_bench_fma_n200_%s.json" % m).read().strip().splitlines()[-1])
        print("FMA build", m, "step %.3f ms kernel %.3f ms value %.4g parity %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["value"], d.get("parity")))
    except Exception as ex:
        print("FMA build", m, "failed", ex)
PY
bash scripts/ncu_traffic.sh $T; cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
timeout 900 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; echo rc=$?; cut -c1-300 gpurun_out/${T}_bench_n400.json
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1; tail -6 gpurun_out/${T}_pytest.log
