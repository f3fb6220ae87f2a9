#!/bin/bash
# default bench line, the reference arm and the 1M-group workloads; summaries on stdout, JSON lines under gpurun_out/
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err | cut -c1-300
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
for w in 1m1b cfg3; do timeout 300 python bench.py --workload $w --skip-cpu > gpurun_out/bench_$w.json 2>/dev/null; done
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("cfg2", d["value"] / 1e9, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline_accept"]["frac"])
l = d["roofline_1m_groups"]
print("1m ctx", l["k_round"], l["k_accept"]["frac"])
e = d["e2e"]
print("e2e", e["value"] / 1e9, e["ms_per_step"], e["p50_decide_latency_ms"], e["sync_full"]["value"] / 1e9)
print("cpu", d["cpu_baseline"]["value"] / 1e6, d["clocks"], d["gpu_launches"], d["p50_decide_latency_ms"])
r = json.load(open("gpurun_out/bench_ref.json"))
print("ref", r["value"] / 1e6, r["cpu_baseline"]["cores"])
for w in ("1m1b", "cfg3"):
    d = json.load(open("gpurun_out/bench_%s.json" % w))
    print(w, d["value"] / 1e9, d["ms_per_step"], d["roofline"]["frac"], d["roofline_accept"]["frac"], d["e2e"]["value"] / 1e9,
          d["roofline_accept"]["phase_pipeline_decisions_per_sec"] / 1e9)
PY
