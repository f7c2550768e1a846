# scripts/gpu_r02y.sh — contact pair kernel: resident CTAs per SM (register cap) A/B, one B200
set -x
mkdir -p gpurun_out
for V in base cmb2 cmb4; do
  LIB=nimblesm_b200/lib/variants/libnsm_b200_$V.so
  [ $V = base ] && LIB=nimblesm_b200/lib/libnsm_b200.so
  NSM_B200_LIB=$LIB timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02y_launches_contact_$V.csv python bench.py --workload contact --n 200 --steps 3 --warmup 3 > gpurun_out/r02y_ncu_$V.log 2>&1
  echo "VARIANT $V"; grep contact_pair gpurun_out/r02y_launches_contact_$V.csv | tail -2 | awk -F'","' '{print $NF}'
done
