# scripts/gpu_r02n.sh — contact kernels re-measured (one B200): parity tests, the contact workload, its launch list
set -x
T=${1:-r02n}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_contact.py -q -m gpu -x -k "force_vs_oracle or steps_vs_oracle or cubes_contact" 2>&1 | tail -3
timeout 600 python bench.py --workload contact --n 200 --steps 20 > gpurun_out/${T}_bench_contact_n200.json 2> gpurun_out/${T}_bench_contact_n200.err; tail -2 gpurun_out/${T}_bench_contact_n200.err; python -c "
import json,sys; d=json.load(open('gpurun_out/${T}_bench_contact_n200.json')); print(d['value'], d['contact'], d['parity']['max_rel_fc'], d['parity']['ok'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_contact_n200.csv python bench.py --workload contact --n 200 --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launch_contact.log 2>&1
grep contact_ gpurun_out/${T}_launches_contact_n200.csv | tail -4 | cut -c1-200
