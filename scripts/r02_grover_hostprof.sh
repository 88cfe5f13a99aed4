#!/bin/bash
mkdir -p gpurun_out
QCC_B200_TRACE_FLUSH=1 python - <<'PY' 2>&1 | tail -60 | tee gpurun_out/r02_grover_hostprof.log
import sys, time, cProfile, pstats
sys.path.insert(0, '.')
import numpy as np
t00 = time.perf_counter()
from qcc_b200 import _cabi, circuit, workloads
np.random.seed(0)
t0 = time.perf_counter()
print("import s", t0 - t00)
pr = cProfile.Profile(); pr.enable()
qc, bits = workloads.grover_circuit(14)
t1 = time.perf_counter()
print("gates issued s", t1 - t0)
mb, mp = qc.psi.maxprob()
qc.sync()
t2 = time.perf_counter()
pr.disable()
print("readout+sync s", t2 - t1, mp)
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
PY
