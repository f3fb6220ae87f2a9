#!/usr/bin/env python
"""DRAM traffic of the dominant kernels, measured with ncu on the bench's own command (under gpurun):

    python tools/measure_traffic.py            # writes profiles/traffic.json

bench.py reads profiles/traffic.json for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch,
B200_PROFILING.md) instead of carrying numbers in its source; the file records which command produced each entry.
A number taken under ncu is never a bench value -- only the byte counts are used."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "traffic.json")
RUNS = [
    ("cfg2", ["--skip-large"], ["k_round", "k_accept", "k_tally_slots", "k_commit"]),
    ("1m1b", ["--workload", "1m1b"], ["k_round", "k_accept", "k_tally_slots", "k_commit"]),
]


def main():
    res = {}
    for name, extra, kernels in RUNS:
        log = os.path.join(ROOT, "gpurun_out", f"traffic_{name}.csv")
        os.makedirs(os.path.dirname(log), exist_ok=True)
        cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control",
               "none", "-k", "regex:^(" + "|".join(kernels) + ")$", "-c", "40", "--csv", "--log-file", log, sys.executable,
               os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--skip-cpu", "--skip-e2e"] + extra
        subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
        rows = [r for r in csv.reader(open(log)) if len(r) > 10 and r[0].isdigit()]
        per = {}
        for r in rows:
            k = r[4].split("(")[0].replace("void ", "").split("<")[0]
            metric, unit, val = r[-3], r[-2], float(r[-1].replace(",", ""))
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3}.get(unit, 1)
            per.setdefault((k, r[0]), {})[metric] = val * scale
        by_k = {}
        for (k, _), m in per.items():
            by_k.setdefault(k, []).append(m)
        for k, ms in by_k.items():
            ms = ms[len(ms) // 2:]  # the later launches (steady state)
            rd = sum(m.get("dram__bytes_read.sum", 0) for m in ms) / len(ms)
            wr = sum(m.get("dram__bytes_write.sum", 0) for m in ms) / len(ms)
            res[f"{name}:{k}"] = {"dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "traffic": int(rd + wr),
                                  "ncu_time_us": sum(m.get("gpu__time_duration.sum", 0) for m in ms) / len(ms) / 1e3,
                                  "launches_averaged": len(ms), "command": " ".join(cmd[cmd.index(sys.executable) + 1:])}
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "traffic.json"), "w"), indent=1, sort_keys=True)  # travels back
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
