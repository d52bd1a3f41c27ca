#!/bin/bash
for b in 0 1; do
  echo "builder=$b : "
  FS_TUNE_BUILDER=$b timeout 200 python tools/profile_step.py 1 2 concert_hall 32 | tail -2 | cut -c1-330
done
