# scripts/gpu_r02r.sh — A/B of the CTA shape of the element kernel: 4 x 128 threads and 1 x 512 threads per SM against 2 x 256.
set -x
T=r02r
mkdir -p gpurun_out
for V in base t128b4 t512b1; do
  for MAT in elastic neohookean; do
    LIB=nimblesm_b200/lib/variants/libnsm_b200_$V.so
    [ $V = base ] && LIB=nimblesm_b200/lib/libnsm_b200.so
    NSM_B200_LIB=$LIB timeout 300 python bench.py --n 200 --material $MAT --steps 20 --no-e2e --no-cpu --no-parity \
      > gpurun_out/${T}_variant_${V}_${MAT}.json 2> gpurun_out/${T}_variant_${V}_${MAT}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_variant_${V}_${MAT}.json").read().strip().splitlines()[-1])
    print("VARIANT %-10s %-14s step %.3f ms  elem kernel %.3f ms  fp64 frac %.4f  clocks %s" % ("$V", "$MAT", d["ms_per_step"], d["roofline"]["kernel_ms"], d["fp64"]["frac"], d["clocks"]))
except Exception as e:
    print("VARIANT $V $MAT failed", e, open("gpurun_out/${T}_variant_${V}_${MAT}.err").read()[-500:])
PY
  done
done
