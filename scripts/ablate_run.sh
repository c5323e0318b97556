#!/bin/bash
# diagnostics: bench.py (short sweep) under each ablation library in build_tmp/ and under fewer resident CTAs per SM
run() { python bench.py --steps 3 --warmup 2 --T 200 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/sweep %.3f' % d['ms_per_step'], 'kernel us %.2f' % (1e3 * d['roofline']['kernel_ms_per_launch']))
"; }
run base
for n in ; do STOCHY_B200_LIB=$PWD/build_tmp/lib_ablate$n.so run ablate$n; done
for c in ; do SB_CTAS_PER_SM=$c run ctas$c; done
