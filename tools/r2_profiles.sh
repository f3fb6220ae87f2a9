#!/bin/bash
# round-2 evidence: traffic.json, launch lists and ncu --set full captures of the final kernels (1 GPU, under gpurun)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e"
python tools/measure_traffic.py > gpurun_out/traffic_stdout.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2g_launches_cfg2.csv $B --skip-large > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_round$ -s 4 -c 2 -o gpurun_out/prof_round_cfg2_r2g -f $B --skip-large > gpurun_out/ncu_r2g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_round$ -s 4 -c 2 -o gpurun_out/prof_round_1m_r2g -f $B --workload 1m1b >> gpurun_out/ncu_r2g.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^(k_accept|k_tally_slots|k_commit)$" -s 9 -c 3 -o gpurun_out/prof_phases_1m_r2g -f $B --workload 1m1b >> gpurun_out/ncu_r2g.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2g_launches_spread_local.csv python bench.py --placement spread --skip-cpu --skip-e2e --steps 2 --warmup 1 --no-graph > /dev/null 2>&1
ls -la gpurun_out | grep r2g
