#!/bin/bash
# diagnostics: scripts/cfg3_split.py under the built library and under each variant in build_tmp/ (three rounds, interleaved)
for rep in 1 2 3; do
  echo base $(python scripts/cfg3_split.py 2>/dev/null)
  for f in build_tmp/*.so; do echo $(basename $f .so) $(STOCHY_B200_LIB=$PWD/$f python scripts/cfg3_split.py 2>/dev/null); done
done
