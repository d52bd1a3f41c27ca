#!/bin/bash
export PS_NOTIME=1
run() { echo "== $1"; python tools/profile_step.py 0 4 $2 $3 2>&1 | grep "step 3" | cut -c1-60; }
for v in "" t128 t64; do
  if [ -n "$v" ]; then export FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/$v/libfrequensee.so; fi
  for s in 1 2 3; do FS_TUNE_STREAMS=$s run "room threads=[$v] streams=$s" furnished_room 16; done
done
export PS_PATHS=1310720
for v in t128 t64; do
  export FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/$v/libfrequensee.so
  for s in 2 3; do FS_TUNE_STREAMS=$s run "hall 1.31M threads=[$v] streams=$s" concert_hall 32; done
done
