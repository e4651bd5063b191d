#!/bin/bash
# round 2, session S (1 GPU, the last 2 GPU-minutes): smoke() and a slice of the GPU suite at HEAD (planner portfolio, traffic-based
# plan choice, four-entry plan cache, store-side variant compiled out of the plain kernels)
mkdir -p gpurun_out
T0=$SECONDS
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 75 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "plan_cache or support_tracking or random_circuits or known_answers or fused_equals_unfused" > gpurun_out/r2s_pytest.log 2>&1
echo "pytest exit $? ($((SECONDS-T0)) s)"; tail -3 gpurun_out/r2s_pytest.log | cut -c1-300
