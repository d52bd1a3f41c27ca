#!/bin/bash
# compute-sanitizer over tools/sanitize_workload.py (run under gpurun): memcheck (out-of-bounds / misaligned accesses, leaks of
# the CUDA API) and racecheck (shared-memory hazards: the warp-shared triangle queue, the shade queue, the FFT buffers).
# The summaries go to gpurun_out/<tag>_sanitizer_{memcheck,racecheck}.log  ->  profiles/.
T=${1:-r2}
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_workload.py > gpurun_out/${T}_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/${T}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python tools/sanitize_workload.py > gpurun_out/${T}_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/${T}_sanitizer_racecheck.log
tail -4 gpurun_out/${T}_sanitizer_memcheck.log; tail -4 gpurun_out/${T}_sanitizer_racecheck.log
