#!/bin/bash
run() { echo "== $1"; python tools/profile_step.py 0 3 $2 $3 2>&1 | grep "step 2" | cut -c1-130; }
run "room default" furnished_room 16
FS_LIB_PATH=$PWD/audio-pathtracer_b200/lib_var/sh8/libfrequensee.so run "room shade 8 CTAs/SM" furnished_room 16
