#!/bin/bash
# diagnostics: programmatic-dependent-launch edges on sharded pools (SB_PDL_DIST bit 0: scan may be scheduled while the
# step kernel drains; bit 1: step kernel's prologue overlaps the scan), N ranks on N GPUs.   usage: pdl_dist_sweep.sh N
N=${1:-2}
for rep in 1 2; do for v in 2 3 0 1; do
  SB_PDL_DIST=$v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$v bench.py --gpus $N --steps 4 --warmup 2 --T 250 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('SB_PDL_DIST=$v ms/sweep %.3f' % d['ms_per_step'], 'kernel us %.2f' % (1e3 * d['roofline']['kernel_ms_per_launch']), 'parity', d['parity_vs_single_gpu']['ok'])
"
done; done
