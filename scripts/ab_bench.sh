#!/bin/bash
# diagnostics: bench.py (short sweeps; extra args pass through, e.g. --workload lg --precision fp64) under the built library
# and under each variant library in build_tmp/ (three rounds, interleaved: compare within one box)
run() { python bench.py --steps 4 --warmup 2 --T 250 "${@:2}" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'ms/sweep %.3f' % d['ms_per_step'], 'kernel us %.2f' % (1e3 * d['roofline']['kernel_ms_per_launch']), d['check']['particles_blake2b64'])
"; }
for rep in 1 2 3; do
  run built "$@"
  for f in build_tmp/*.so; do STOCHY_B200_LIB=$PWD/$f run $(basename $f .so) "$@"; done
done
