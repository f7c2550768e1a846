# scripts/gpu_final.sh [tag] — the round's closing measurements on one B200 with the final library: smoke, the GPU test
# suite, the default bench line (64 M neohookean) with the reference arm beside it, ORDERED / two-block / configs[1] /
# contact lines, the ncu launch lists of the default and the contact commands, stamped ncu traffic of the element
# kernels, compute-sanitizer on the smoke run and on the tests that exercise the newest kernels.
set -x
T=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench_n400.json 2> gpurun_out/${T}_bench_n400.err; tail -3 gpurun_out/${T}_bench_n400.err; cut -c1-300 gpurun_out/${T}_bench_n400.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1; cut -c1-300 gpurun_out/${T}_bench_ref.json
timeout 600 python bench.py --assembly ordered --steps 10 --no-cpu > gpurun_out/${T}_bench_n400_ordered.json 2>&1; cut -c1-200 gpurun_out/${T}_bench_n400_ordered.json
timeout 600 python bench.py --workload twoblock --steps 10 --no-cpu > gpurun_out/${T}_bench_n400_twoblock.json 2>&1; cut -c1-200 gpurun_out/${T}_bench_n400_twoblock.json
timeout 600 python bench.py --workload contact --n 200 --steps 20 > gpurun_out/${T}_bench_contact_n200.json 2> gpurun_out/${T}_bench_contact.err; cut -c1-200 gpurun_out/${T}_bench_contact_n200.json
timeout 600 python bench.py --workload contact --n 320 --steps 10 > gpurun_out/${T}_bench_contact_n320.json 2>> gpurun_out/${T}_bench_contact.err; cut -c1-200 gpurun_out/${T}_bench_contact_n320.json
bash scripts/bench_config1.sh ${T}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_n400.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_contact_n200.csv python bench.py --workload contact --n 200 --steps 3 --warmup 3 > gpurun_out/${T}_ncu_launch_contact.log 2>&1
bash scripts/ncu_traffic.sh ${T} > gpurun_out/${T}_ncu_traffic.log 2>&1; tail -1 gpurun_out/${T}_ncu_traffic.log | cut -c1-300
bash scripts/gpu_sanitize.sh ${T} > gpurun_out/${T}_sanitize.log 2>&1; cat gpurun_out/${T}_compute_sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_contact.py tests/test_gpu_parity.py -m gpu -q -x -k "force_vs_oracle or steps_vs_oracle or argument or never_misses or pipelined_internal_force" > gpurun_out/${T}_san_memcheck_new.log 2>&1; echo "memcheck(new) rc=$?"; tail -4 gpurun_out/${T}_san_memcheck_new.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_contact.py -m gpu -q -x -k "force_vs_oracle" > gpurun_out/${T}_san_racecheck_contact.log 2>&1; echo "racecheck(contact) rc=$?"; tail -4 gpurun_out/${T}_san_racecheck_contact.log
ls -la gpurun_out | tail -5
