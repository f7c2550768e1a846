# scripts/gpu_r02b.sh — round 2, second GPU call: the GPU suite (state-variable slot, lockstep multi-rank fix) and
# element-kernel A/B variants (occupancy / register cap / b^-1 staging / lazy sC) on the 8 M-element cube.
set -x
T=r02b
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${T}_pytest.log 2>&1; tail -25 gpurun_out/${T}_pytest.log
for V in base lazysc nostage w18r112 w18r112ns; do
  for MAT in neohookean elastic; do
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
ls gpurun_out | tail -30
