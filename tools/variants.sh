#!/bin/bash
# usage: tools/variants.sh "<variant names under build/variants/ | base>" "<workloads>"  -> gpurun_out/variants.txt
# each variant is an alternative build of the same CUDA library (python -m gigapaxos_b200.build with -D tuning
# defines); GPX_LIB selects it.
out=gpurun_out/variants.txt; mkdir -p gpurun_out; : > $out
for v in $1; do for w in $2; do
  lib=$PWD/build/variants/libgpx_$v.so; [ "$v" = base ] && lib=$PWD/gigapaxos_b200/libgpx.so
  GPX_LIB=$lib python bench.py --workload $w --skip-cpu --skip-e2e --skip-large --steps 40 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d.get('roofline') or {}
print('$v','$w',round(d['value']/1e9,3),'Gdec/s',round(d['ms_per_step']*1e3,2),'us/step frac',round(r.get('frac',0),4),'k_round_us',round(r.get('kernel_ms',0)*1e3,2))" >> $out 2>&1
done; done
cat $out
