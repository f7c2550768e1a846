#!/bin/bash
# sanitizer passes over the final pair kernel (shared candidate list), then the contact line again
T=r02W
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_contact.py -m gpu -q -x -k "force_vs_oracle" > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/${T}_racecheck.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_contact.py -m gpu -q -x -k "force_vs_oracle or never_misses" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${T}_memcheck.log
timeout 300 python bench.py --workload contact --n 200 --steps 20 > gpurun_out/${T}_bench_contact_n200.json 2> gpurun_out/${T}_bench_contact_n200.err; cut -c1-260 gpurun_out/${T}_bench_contact_n200.json
