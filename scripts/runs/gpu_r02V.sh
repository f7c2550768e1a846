#!/bin/bash
# pair kernel with the chain walk separated from the box tests: contact suite, bench line, launch list
T=r02V
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_contact.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload contact --n 200 --steps 20 > gpurun_out/${T}_bench_contact_n200.json 2> gpurun_out/${T}_bench_contact_n200.err; cut -c1-300 gpurun_out/${T}_bench_contact_n200.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches_contact_n200.csv python bench.py --workload contact --n 200 --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launch_contact.log 2>&1
grep -E "contact_(pair|bin|update)" gpurun_out/${T}_launches_contact_n200.csv | tail -6 | cut -d, -f5,15-
