cd ${GRAFT_REPO_ROOT:-.}
P=/tmp/prof; mkdir -p $P gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
name=r2_step_hmm_peers
timeout 300 $NCU -k regex:k_step_hmm -s 20 -c 1 -o $P/$name python bench.py --steps 2 --warmup 3 --T 40 --force-shard > gpurun_out/ncu_${name}.log 2>&1
python scripts/ncu_summary.py full $P/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
ncu -i $P/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
name=r2_tilescan_peers
timeout 300 $NCU -k regex:k_tilescan -s 20 -c 1 -o $P/$name python bench.py --steps 2 --warmup 3 --T 40 --force-shard > gpurun_out/ncu_${name}.log 2>&1
python scripts/ncu_summary.py full $P/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
