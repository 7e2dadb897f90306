#!/bin/bash
# One GPU-box session: parity tests, smoke, pipe peaks, bench.  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 120 tools/pipe_peaks > gpurun_out/pipe_peaks.json 2> gpurun_out/pipe_peaks.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
for t in chamfer knn emd; do
  timeout 900 python -m pytest tests/test_gpu_$t.py -q -m gpu --maxfail=6 -x --tb=short > gpurun_out/test_$t.log 2>&1; echo "rc=$?" >> gpurun_out/test_$t.log
done
timeout 900 python -m pytest tests/test_ref_cuda_parity.py -q -m gpu --tb=short > gpurun_out/test_ref.log 2>&1; echo "rc=$?" >> gpurun_out/test_ref.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -n 5 gpurun_out/smoke.log gpurun_out/test_*.log gpurun_out/bench.err
cat gpurun_out/pipe_peaks.json
