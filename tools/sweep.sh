#!/bin/bash
cp audio-pathtracer_b200/lib/libfrequensee.so /tmp/orig.so
for v in g0 g0mb5; do
  cp audio-pathtracer_b200/lib/var_$v.bin audio-pathtracer_b200/lib/libfrequensee.so
  echo -n "$v : "
  timeout 60 python tools/profile_step.py 0 3 | tail -1 | cut -c1-130
done
cp /tmp/orig.so audio-pathtracer_b200/lib/libfrequensee.so
