#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, then the ncu launch list of the same bench command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -q -m gpu --maxfail=6 --tb=short > gpurun_out/test_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/test_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; rc=$?; echo "bench rc=$rc" >> gpurun_out/bench.err
if [ "$1" = "ncu" ] && [ $rc -eq 0 ]; then
  timeout 600 python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 2 --warmup 3 --no-sub --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?" >> gpurun_out/bench.err
fi
tail -n 4 gpurun_out/smoke.log gpurun_out/test_gpu.log gpurun_out/bench.err
