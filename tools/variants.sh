#!/bin/bash
# usage: scripts_variants.sh "<variants>" "<workloads>"  -> gpurun_out/variants.txt
out=gpurun_out/variants.txt; mkdir -p gpurun_out; : > $out
for v in $1; do for w in $2; do
  lib=/root/repo/gigapaxos_b200/libgpx$v.so; [ "$v" = base ] && lib=/root/repo/gigapaxos_b200/libgpx.so
  GPX_LIB=$lib python bench.py --workload $w --skip-cpu --skip-e2e --skip-large --steps 30 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
print('$v','$w',round(d['value']/1e9,3),'Gdec/s',d['ms_per_step'],'ms frac',r.get('frac'), d.get('kernel_us'))" >> $out 2>&1
done; done
cat $out
