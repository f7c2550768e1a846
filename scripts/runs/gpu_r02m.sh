# scripts/gpu_r02m.sh — contact measurements on one B200: the contact workload of bench.py (8 M elements in two stacked
# bodies) in both assembly modes, a smaller one, the ncu launch list of contact steps, compute-sanitizer on the contact tests.
set -x
T=${1:-r02m}
mkdir -p gpurun_out
timeout 600 python bench.py --workload contact --n 200 --steps 20 > gpurun_out/${T}_bench_contact_n200.json 2> gpurun_out/${T}_bench_contact_n200.err; tail -2 gpurun_out/${T}_bench_contact_n200.err; cut -c1-1800 gpurun_out/${T}_bench_contact_n200.json
timeout 600 python bench.py --workload contact --n 200 --steps 20 --assembly ordered > gpurun_out/${T}_bench_contact_n200_ordered.json 2>> gpurun_out/${T}_bench_contact_n200.err; cut -c1-300 gpurun_out/${T}_bench_contact_n200_ordered.json
timeout 600 python bench.py --workload contact --n 320 --steps 10 > gpurun_out/${T}_bench_contact_n320.json 2>> gpurun_out/${T}_bench_contact_n200.err; cut -c1-300 gpurun_out/${T}_bench_contact_n320.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_contact_n200.csv python bench.py --workload contact --n 200 --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launch_contact.log 2>&1
grep -c contact gpurun_out/${T}_launches_contact_n200.csv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_contact.py -m gpu -q -x -k "force_vs_oracle or steps_vs_oracle or argument" > gpurun_out/${T}_san_memcheck_contact.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_san_memcheck_contact.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_contact.py -m gpu -q -x -k "force_vs_oracle" > gpurun_out/${T}_san_racecheck_contact.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/${T}_san_racecheck_contact.log
