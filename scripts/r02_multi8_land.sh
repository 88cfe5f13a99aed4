#!/bin/bash
# r02, 8 GPUs: landing bits (arriving qubits on the highest local bits) against plain swaps on the large shards.
set -u
N=8
mkdir -p gpurun_out
trun() {
  name=$1; shift
  PORT=$((29300 + RANDOM % 500))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N "$@" 2> gpurun_out/r02_$name.err | grep '^{' > gpurun_out/r02_$name.json
  echo "== $name: $(wc -l < gpurun_out/r02_$name.json) line(s)"
}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "push-8" 2>&1 | tail -3 | tee gpurun_out/r02_pytest_multi8_land.log
QCC_B200_TRACE_FLUSH=1 trun qft34_8gpu_land --qubits 34 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e
grep "qcc_b200 launch" gpurun_out/r02_qft34_8gpu_land.err | sort | uniq -c | sort -k3 -n | tail -40 > gpurun_out/r02_qft34_8gpu_land_launches.txt
QCC_B200_NO_LAND=1 trun qft34_8gpu_noland --qubits 34 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e
trun scale_qft30_8_land --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
trun supremacy34_8gpu_land --workload supremacy --qubits 34 --depth 20
trun matrix_8_land --matrix supremacy:34,supremacy:32,qft:32 --steps 4 --warmup 3
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_*land*.json")):
  for ln in open(f):
    try:
      d=json.loads(ln)
    except Exception:
      continue
    cfg=d.get("config",{})
    print(f.split("/")[-1], cfg.get("workload", d.get("workload")), cfg.get("qubits", d.get("qubits")), "ms/step %.2f"%d.get("ms_per_step",-1),
          "passes", d.get("passes_per_step", d.get("passes")), "kernel_ms", d.get("kernel_ms"), "nvlink", (d.get("exchange") or {}).get("nvlink_gbs_per_direction_rank0"), "check", (d.get("check") or {}).get("ok"), (d.get("parity_vs_single_gpu") or {}).get("ok"), d.get("error"))
PY
tail -12 gpurun_out/r02_qft34_8gpu_land_launches.txt | cut -c1-200
