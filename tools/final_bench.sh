#!/bin/bash
# round-2 bench lines on ONE GPU (under gpurun): default (cfg2), the reference arm, 1 M groups, cfg4, cfg5, spread with
# four nodes on one GPU; JSON lines under gpurun_out/, a one-line summary each on stdout.  Multi-GPU lines:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29500 \
#       bench.py --gpus N [--workload cfg3] [--p2p] [--placement packed]
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_{n}.json"))
    e = d.get("e2e") or {}
    print(n, "value %.3f G" % (d["value"] / 1e9), "ms %.4f" % d["ms_per_step"], "frac %.3f" % (d.get("roofline") or {}).get("frac", 0),
          "e2e %.3f G" % (e.get("value", 0) / 1e9))
except Exception as ex:
    print(n, "FAILED", ex)
PY
}
run default
run ref --impl reference
run 1m1b --workload 1m1b --skip-cpu
run cfg4 --workload cfg4 --skip-cpu --steps 10
run cfg5 --workload cfg5 --skip-cpu --steps 10
run spread_local --placement spread --skip-cpu --steps 10
