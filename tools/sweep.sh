#!/bin/bash
for w in 1 0; do
  echo -n "wide=$w : "
  FS_TUNE_WIDE=$w timeout 100 python tools/profile_step.py 1 3 | tail -1 | cut -c1-330
  echo -n "wide=$w nocount: "
  FS_TUNE_WIDE=$w timeout 100 python tools/profile_step.py 0 3 | tail -1 | cut -c1-130
done
