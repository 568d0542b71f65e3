#!/bin/bash
# A/B of experiment builds (scratch_libs/libb200rt_<name>.so) against the product library: scripts/variants.sh spp name...
spp=$1; shift
summ() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['cfg'], d['Mpaths/s'], '(%d Mrays)' % d['Mrays/s'], end='  ')
print()"; }
echo -n "product   "; python scripts/perf_all.py $spp 2>&1 | summ
for v in "$@"; do echo -n "$v   "; B200RT_LIB=$PWD/scratch_libs/libb200rt_$v.so python scripts/perf_all.py $spp 2>&1 | summ; done
