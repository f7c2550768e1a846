set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/san_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or multi_step or bc_programs or element_components" > gpurun_out/san_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -6 gpurun_out/san_memcheck_tests.log
