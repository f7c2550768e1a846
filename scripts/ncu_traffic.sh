# scripts/ncu_traffic.sh [tag] — one `ncu --set full` capture of the element kernel per material (200^3 cube, the
# cached-Jacobian instance bench.py runs), summaries under gpurun_out/, and profiles/ncu_traffic.json (DRAM bytes per
# element stamped with the kernel-source hash; bench.py reports it as roofline.traffic while the hash matches).
T=${1:-r02}
mkdir -p gpurun_out
for MAT in neohookean elastic; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:element_force -s 3 -c 1 -f \
    -o gpurun_out/${T}_ncu_elem_${MAT}_f2 python bench.py --n 200 --material $MAT --steps 1 --no-e2e --no-cpu --no-parity --flags 2 \
    > gpurun_out/${T}_ncu_elem_${MAT}.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${T}_ncu_elem_${MAT}_f2.ncu-rep > gpurun_out/${T}_ncu_elem_${MAT}_f2_summary.txt 2>&1
  python scripts/ncu_sass_mix.py gpurun_out/${T}_ncu_elem_${MAT}_f2.ncu-rep 2000000 >> gpurun_out/${T}_ncu_elem_${MAT}_f2_summary.txt 2>&1
done
python scripts/ncu_traffic.py gpurun_out/ncu_traffic.json \
  mat1_ordered0_mode2=gpurun_out/${T}_ncu_elem_neohookean_f2.ncu-rep:8000000 \
  mat0_ordered0_mode2=gpurun_out/${T}_ncu_elem_elastic_f2.ncu-rep:8000000
