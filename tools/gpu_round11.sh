#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_orbit.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | tail -2
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/profile_case.py c4s 1 > gpurun_out/sanitizer_racecheck_orbit.log 2>&1; echo "racecheck c4s rc=$?"; tail -1 gpurun_out/sanitizer_racecheck_orbit.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/profile_case.py c4e 1 > gpurun_out/sanitizer_racecheck_orbit_c4e.log 2>&1; echo "racecheck c4e rc=$?"; tail -1 gpurun_out/sanitizer_racecheck_orbit_c4e.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/profile_case.py c4s 1 > gpurun_out/sanitizer_synccheck_orbit.log 2>&1; echo "synccheck rc=$?"; tail -1 gpurun_out/sanitizer_synccheck_orbit.log
timeout 300 python tools/exp_orbit.py 20 c4,c4p 2>&1 | grep -E "us=|ERROR" | head -3
