#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, ncu launch list.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
echo "== pytest -m gpu" | tee gpurun_out/pytest_gpu.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee -a gpurun_out/pytest_gpu.log
echo "== smoke" | tee gpurun_out/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --workload hsweep30 --steps 3 --no-e2e --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_hsweep30.json
timeout 600 python bench.py --workload larose28 --steps 3 --no-e2e --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_larose28.json
timeout 900 python bench.py --steps 10 2>&1 | tail -2 | tee gpurun_out/bench_qft30.json
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_qft30.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
    > gpurun_out/ncu_qft30.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_hsweep30.csv python bench.py --workload hsweep30 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline \
    > gpurun_out/ncu_hsweep30.log 2>&1
fi
echo done
