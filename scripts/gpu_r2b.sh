#!/bin/bash
# round 2, session B (1 GPU): I/O microbenchmark + ncu of the classic2 and ring forms of the dense qft30 passes
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 ./scripts/microbench3 30 > gpurun_out/r2b_microbench3.log 2>&1
cat gpurun_out/r2b_microbench3.log
for form in classic2 ring; do
  DVD_JIT_FORM=$form timeout 600 ncu --set full --clock-control none --import-source on -k regex:dvd_pass_static -s 9 -c 3 \
      -o gpurun_out/r2b_ncu_qft30_$form -f python bench.py --workload qft30 --jit 1 --steps 1 --warmup 1 --no-cpu-baseline --no-scaling-point --no-e2e --no-single-gate > gpurun_out/r2b_ncu_$form.log 2>&1
  tail -3 gpurun_out/r2b_ncu_$form.log | cut -c1-300
done
echo "total $((SECONDS-T0)) s"
