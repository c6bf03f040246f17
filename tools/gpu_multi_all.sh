#!/bin/bash
TAG=${1:-r3}
# 8-GPU box: BASELINE.json configs[2] (AAS step), configs[3] (paired, 2/4/8) and configs[4] (sweep) under torchrun
mkdir -p gpurun_out/scaling
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc port args...
  name=$1; np=$2; port=$3; shift 3
  if [ "$np" = "1" ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scaling/$name.json 2> gpurun_out/scaling/$name.err
  else timeout 900 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > gpurun_out/scaling/$name.json 2> gpurun_out/scaling/$name.err; fi
  python - gpurun_out/scaling/$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline",{}); print("%-22s n_gpus %d value %.4e ms/step %.4f e2e %.4e step_frac %s" % (sys.argv[2], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("step_frac")))
    if "aas_step" in d: print("   aas_step:", {k:v for k,v in d["aas_step"].items() if k!="models"})
except Exception as e:
    print(sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
}
nvidia-smi -L | head -8
echo "== CPU column of the sweep (one process)"; timeout 600 python tools/sweep.py --cpu-only --quick 2>&1 | tail -4
run ${TAG}_chime_n1 1 0 --steps 20 --warmup 5 --no-cpu --no-large
for np in 2 4 8; do run ${TAG}_chime_n$np $np 2960$np --steps 20 --warmup 5 --no-cpu --no-large; done
run ${TAG}_paired_n1 1 0 --workload paired_30x6s --steps 20 --warmup 5 --no-large --no-cpu
for np in 2 4 8; do run ${TAG}_paired_n$np $np 2961$np --workload paired_30x6s --steps 20 --warmup 5 --no-large --no-cpu; done
run ${TAG}_sweep256_n1 1 0 --workload sweep_256x10s --steps 20 --warmup 5 --no-cpu --no-e2e
run ${TAG}_sweep256_n8 8 29621 --workload sweep_256x10s --steps 20 --warmup 5 --no-cpu --no-e2e
run ${TAG}_aas_n1 1 0 --workload aas_step_30x6s --steps 6 --warmup 3
run ${TAG}_aas_n8 8 29631 --workload aas_step_30x6s --steps 6 --warmup 3
echo "== host-fabric ceiling"
timeout 300 python tools/pcie_ceiling.py | tee gpurun_out/scaling/${TAG}_pcie_n1.json
timeout 300 $TR --nproc-per-node 8 --master-port 29651 tools/pcie_ceiling.py 2>/dev/null | tail -1 | tee gpurun_out/scaling/${TAG}_pcie_n8.json
echo "== sweep quick at 8 GPUs"
timeout 900 $TR --nproc-per-node 8 --master-port 29641 tools/sweep.py gpurun_out/scaling/${TAG}_sweep_n8.md --quick 2>&1 | tail -14
