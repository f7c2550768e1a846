# scripts/gpu_r02K.sh — elastic element kernel: CTAs of 64 and 32 threads against the 128 of the default, one B200
set -x
mkdir -p gpurun_out
for V in base e64 e32; do
  LIB=nimblesm_b200/lib/variants/libnsm_b200_$V.so
  [ $V = base ] && LIB=nimblesm_b200/lib/libnsm_b200.so
  NSM_B200_LIB=$LIB timeout 300 python bench.py --n 200 --material elastic --steps 20 --no-e2e --no-cpu --no-parity > gpurun_out/r02K_variant_${V}_elastic.json 2> gpurun_out/r02K_variant_${V}_elastic.err
  python -c "
import json; d=json.load(open('gpurun_out/r02K_variant_${V}_elastic.json')); print('VARIANT %-8s elastic step %.3f ms  elem kernel %.3f ms  fp64 frac %.4f  clocks %s' % ('$V', d['ms_per_step'], d['roofline']['kernel_ms'], d['fp64']['frac'], d['clocks']))"
done
