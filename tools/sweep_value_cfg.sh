#!/bin/bash
# value-run configurations: reads/s against per-launch efficiency of the extension kernel (run after bench.py filled its cache)
D=/tmp/yaha_b200_bench_iid100; X=$(ls $D/ref.X15_01_* | head -1); Q=$D/reads_rank0.fa; H=yaha_b200/yaha_b200_host
for spec in "5000 4 28000" "10000 2 28000" "10000 2 16000" "10000 2 0" "6667 3 28000" "6667 3 12000" "20000 1 28000" "10000 4 16000"; do
  set -- $spec
  YA_PACKED_NARROW_BELOW=$3 $H -x $X -q $Q -osh /tmp/sv.sam -t 4 -batch $1 -pipes $2 -passes 14 -replay -BW 10 -G 100 2>&1 | grep '"pass"' | tail -8 | python3 -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
a=sorted(x['align_s'] for x in r)
c=sum(x['ext_cells'] for x in r); ms=sum(x['dev_ms_ext'] for x in r); n=sum(x['ext_launches'] for x in r); u=sum(x['dev_ms_ext_union'] for x in r)
print('batch %5d pipes %d narrow<%5d: %.2f ms/step %.2f M reads/s | ext: %.0f us/launch, per-launch %.0f GCUPS (%.2f), device-level %.0f GCUPS (%.2f)' % ($1, $2, $3, a[len(a)//2]*1e3, r[0]['reads']/a[len(a)//2]/1e6, 1e3*ms/max(n,1), c/max(ms,1e-9)/1e6, c/max(ms,1e-9)/1e6/1015, c/max(u,1e-9)/1e6, c/max(u,1e-9)/1e6/1015))"
done
