# scripts/gpu_r02P.sh — neohookean element kernel at 17 / 18 single-warp CTAs per SM (120 / 112 registers, small spills)
set -x
mkdir -p gpurun_out
for V in base w17x1 w18x1 base; do
  LIB=nimblesm_b200/lib/variants/libnsm_b200_$V.so
  [ $V = base ] && LIB=nimblesm_b200/lib/libnsm_b200.so
  NSM_B200_LIB=$LIB timeout 300 python bench.py --n 200 --material neohookean --steps 20 --no-e2e --no-cpu > gpurun_out/r02P_variant_${V}_neohookean.json 2> gpurun_out/r02P_variant_${V}_neohookean.err
  python -c "
import json; d=json.load(open('gpurun_out/r02P_variant_${V}_neohookean.json')); print('VARIANT %-8s neohookean step %.3f ms  elem kernel %.3f ms  fp64 frac %.4f  parity %s clocks %s' % ('$V', d['ms_per_step'], d['roofline']['kernel_ms'], d['fp64']['frac'], d['parity']['max_rel_f'], d['clocks']))" || tail -3 gpurun_out/r02P_variant_${V}_neohookean.err
done
