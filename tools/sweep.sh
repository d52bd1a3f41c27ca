#!/bin/bash
for tex in 3 2; do
  echo -n "tex=$tex : "
  FS_TUNE_TEX=$tex timeout 60 python tools/profile_step.py 0 3 | tail -1 | cut -c1-130
done
