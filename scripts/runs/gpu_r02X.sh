#!/bin/bash
# closing check at HEAD: smoke, the whole GPU suite, the default bench line, ncu --set full of the shipped pair kernel
T=r02X
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; echo "bench rc=$?"; cut -c1-220 gpurun_out/${T}_bench_n400.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:contact_pair -s 3 -c 1 -f -o gpurun_out/${T}_ncu_contact_pair \
  python bench.py --workload contact --n 200 --steps 2 --warmup 3 > gpurun_out/${T}_ncu_contact_pair.log 2>&1
python scripts/ncu_summary.py gpurun_out/${T}_ncu_contact_pair.ncu-rep > gpurun_out/${T}_ncu_contact_pair_summary.txt 2>&1; head -30 gpurun_out/${T}_ncu_contact_pair_summary.txt
