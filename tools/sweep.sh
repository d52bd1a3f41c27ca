#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-100; }
for v in tritex tritex2; do
FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/$v/libfrequensee.so run "room $v" furnished_room 16
FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/$v/libfrequensee.so run "hall $v" concert_hall 32
done
