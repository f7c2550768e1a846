set -x
mkdir -p gpurun_out
B="timeout 300 python bench.py --n 200 --steps 10 --no-cpu --no-e2e"
$B > gpurun_out/I_n200_neohookean_lattice.json 2>&1
$B --shuffle > gpurun_out/I_n200_neohookean_shuffled.json 2>&1
$B --shuffle --material elastic > gpurun_out/I_n200_elastic_shuffled.json 2>&1
for f in gpurun_out/I_n200_*.json; do echo $f; python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); print(" value %.4g  elem_ms %.3f node_ms %.3f cold %s"%(d["value"],d["roofline"]["kernel_ms"],d["node_kernels_ms"],d.get("cold_points")))
except Exception as e: print("ERR",e, open("$f").read()[-800:])
PY
done
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:element_force -s 3 -c 1 --csv --log-file gpurun_out/I_shuffled_dram.csv python bench.py --n 200 --steps 1 --no-e2e --no-cpu --shuffle > /dev/null 2>&1
grep -v "^==" gpurun_out/I_shuffled_dram.csv | cut -d, -f5,13,15
