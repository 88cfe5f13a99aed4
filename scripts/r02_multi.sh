#!/bin/bash
# r02 multi-GPU visit: bash scripts/r02_multi.sh N   (gpurun --gpus N, N = 2 | 4 | 8)
#   sharded parity tests at world N under the three exchanges, the default bench line at N, the north-star
#   table (QFT / supremacy.py random circuit at 28-34 qubits) at N, and the BASELINE.json config that belongs
#   to N (4: grover-32; 8: QFT-34 + supremacy-34).
set -u
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_$N.txt 2>&1
trun() {  # name, then bench args
  name=$1; shift
  PORT=$((29300 + RANDOM % 500))
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N "$@" 2> gpurun_out/r02_$name.err | grep '^{' > gpurun_out/r02_$name.json
  echo "== $name: $(wc -l < gpurun_out/r02_$name.json) line(s)"; tail -2 gpurun_out/r02_$name.err | cut -c1-300
}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "push-$N or swap-$N or nccl-$N" 2>&1 | tail -6 | tee gpurun_out/r02_pytest_multi$N.log
trun scale_qft30_$N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
if [ "$N" = 8 ]; then
  QCC_B200_TRACE_FLUSH=1 trun qft34_8gpu --qubits 34 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e
  grep "qcc_b200 launch" gpurun_out/r02_qft34_8gpu.err | sort | uniq -c | sort -k3 -n | tail -40 > gpurun_out/r02_qft34_8gpu_launches.txt
  trun supremacy34_8gpu --workload supremacy --qubits 34 --depth 20
  trun matrix_8 --matrix qft:28,qft:32,supremacy:28,supremacy:30,supremacy:32,supremacy:34 --steps 4 --warmup 3
fi
if [ "$N" = 4 ]; then
  trun grover32_4gpu --workload grover --qubits 32
  trun matrix_4 --matrix qft:28,qft:32,qft:34,supremacy:28,supremacy:30,supremacy:32,supremacy:34 --steps 4 --warmup 3
fi
if [ "$N" = 2 ]; then
  trun matrix_2 --matrix qft:28,qft:32,supremacy:28,supremacy:30,supremacy:32,qft:34 --steps 4 --warmup 3
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_*_$N*.json")+glob.glob("gpurun_out/r02_*${N}gpu.json")):
  for ln in open(f):
    try:
      d=json.loads(ln)
    except Exception:
      continue
    cfg=d.get("config",{})
    print(f.split("/")[-1], cfg.get("workload", d.get("workload")), cfg.get("qubits", d.get("qubits")), "ms/step %.2f"%d.get("ms_per_step",-1), "gates/s %.0f"%d.get("value",-1),
          "passes", d.get("passes_per_step", d.get("passes")), "exch", d.get("exchange"), "check", d.get("check"), d.get("parity_vs_single_gpu"), d.get("error"))
PY
echo done
