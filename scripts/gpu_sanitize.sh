# scripts/gpu_sanitize.sh [tag] — compute-sanitizer memcheck + racecheck of the smoke run, memcheck of the tests that
# exercise tail lanes, multi-step calls, BC programs, the state-variable slot and the pipelined host step.
set -x
T=${1:-r02}
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/${T}_san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/${T}_san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/${T}_san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/${T}_san_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_state.py -m gpu -q -x -k "ragged or multi_step or bc_programs or element_components or pipelined or state_explicit or seam_bitwise" > gpurun_out/${T}_san_memcheck_tests.log 2>&1; echo "memcheck tests rc=$?"; tail -6 gpurun_out/${T}_san_memcheck_tests.log
( echo "== memcheck smoke"; tail -4 gpurun_out/${T}_san_memcheck.log; echo "== racecheck smoke"; tail -4 gpurun_out/${T}_san_racecheck.log; echo "== memcheck tests"; tail -5 gpurun_out/${T}_san_memcheck_tests.log ) > gpurun_out/${T}_compute_sanitizer.txt
