#!/bin/bash
# $1 = number of GPUs: headline + config-4 bench lines, PCIe ceiling, reference arm, NCCL multi tests
N=$1
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
if [ "$N" = "1" ]; then
  python tools/pcie_probe.py 512 > $O/pcie_n1.json 2> $O/pcie_n1.err
  python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
  python bench.py --workload advection_3d --cells 64 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
else
  $TR --master-port 29541 tools/pcie_probe.py 512 > $O/pcie_n$N.json 2> $O/pcie_n$N.err
  $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
  $TR --master-port 29543 bench.py --gpus $N --workload advection_3d --cells 64 --steps 20 --warmup 3 > $O/bench_c4_n$N.json 2> $O/bench_c4_n$N.err
  $TR --master-port 29544 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
  NCCL_DEBUG=WARN $TR --master-port 29545 tests/dist_parity_worker.py > $O/dist_parity_n$N.log 2>&1
  python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/multi_tests_n$N.log 2>&1
fi
for f in $O/pcie_n$N.json $O/bench_n$N.json $O/bench_c4_n$N.json; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$f") if l.startswith("{")][-1])
    keep={k:d.get(k) for k in ("n_gpus","ms_per_step","value","both_gbs_aggregate_per_direction","both_gbs_per_gpu_per_direction")}
    keep["e2e_ms"]=(d.get("e2e") or {}).get("ms_per_step"); keep["checks"]=d.get("checks"); keep["frac"]=(d.get("roofline_rhs") or {}).get("frac")
    print("$f", keep)
except Exception as e:
    print("$f", "ERROR", e); print(open("$f".replace(".json",".err")).read()[-1500:])
PY
done
for f in $O/dist_parity_n$N.log $O/multi_tests_n$N.log; do [ -f $f ] && tail -n 3 $f; done; true
