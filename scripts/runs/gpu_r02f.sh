# scripts/gpu_r02f.sh — round 2, sixth GPU call: issue-mix microbenchmark (fixed timing), b^-1 prefetch variants for the
# elastic kernel, the state-material kernel's throughput (8 M and 64 M elements: the latter exercises the size-aware
# Jacobian cache), configs[1] against the sustained FP64 peak.
set -x
T=r02f
mkdir -p gpurun_out
timeout 300 ./scripts/micro/issue_mix > gpurun_out/${T}_issue_mix.txt 2>&1; cat gpurun_out/${T}_issue_mix.txt
for V in base pfl1 pfl2; do
  for MAT in elastic neohookean; do
    LIB=nimblesm_b200/lib/variants/libnsm_b200_$V.so
    [ $V = base ] && LIB=nimblesm_b200/lib/libnsm_b200.so
    NSM_B200_LIB=$LIB timeout 300 python bench.py --n 200 --material $MAT --steps 20 --no-e2e --no-cpu --no-parity \
      > gpurun_out/${T}_variant_${V}_${MAT}.json 2> gpurun_out/${T}_variant_${V}_${MAT}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${T}_variant_${V}_${MAT}.json"))
    print("VARIANT %-10s %-10s step %.3f ms  elem kernel %.3f ms  fp64 frac %.4f  clocks %s" % ("$V", "$MAT", d["ms_per_step"], d["roofline"]["kernel_ms"], d["fp64"]["frac"], d["clocks"]))
except Exception as e:
    print("VARIANT $V $MAT failed", e, open("gpurun_out/${T}_variant_${V}_${MAT}.err").read()[-500:])
PY
  done
done
for N in 200 400; do
  timeout 900 python bench.py --n $N --material j2_plasticity --steps 10 --no-cpu > gpurun_out/${T}_bench_j2_n$N.json 2> gpurun_out/${T}_bench_j2_n$N.err; echo rc=$?; tail -2 gpurun_out/${T}_bench_j2_n$N.err; cut -c1-250 gpurun_out/${T}_bench_j2_n$N.json
done
bash scripts/bench_config1.sh $T
python - <<'PY'
import json
for f in ("r02f_bench_j2_n200", "r02f_bench_j2_n400", "r02f_bench_config1_8M_elastic_1000steps"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.4g ms %.3f kernel %.3f fp64 %s  device_bytes %.1f GB  clocks %s e2e %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["fp64"], d["device_bytes"] / 1e9, d["clocks"], d.get("e2e", {}).get("value")))
    except Exception as ex:
        print(f, "failed", ex)
PY
