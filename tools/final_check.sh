#!/bin/bash
# Round-end check on one B200: GPU test suite, smoke, then the final evidence script (ncu capture -> traffic stamp -> bench lines).
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -x -q -m gpu ) > gpurun_out/gputests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/gputests.log | head -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
bash tools/profile_run_final.sh
