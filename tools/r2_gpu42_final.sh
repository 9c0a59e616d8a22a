#!/bin/bash
# Round 2 final validation: smoke, the whole GPU suite, then the profile session of tools/r2_gpu38_prof.sh.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_42_smoke.log 2>&1; tail -1 gpurun_out/r2_42_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_42_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_42_pytest.log
tail -4 gpurun_out/r2_42_pytest.log
sed -i 's/T=r2_38/T=r2_42/' tools/r2_gpu38_prof.sh
bash tools/r2_gpu38_prof.sh
