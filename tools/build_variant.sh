#!/bin/bash
# tools/build_variant.sh NAME "-DFS_SSTACK=8 ..." : A/B build of the library into audio-pathtracer_b200/lib_var/NAME/
# (use with FS_LIB_PATH=audio-pathtracer_b200/lib_var/NAME/libfrequensee.so)
set -e
cd "$(dirname "$0")/../audio-pathtracer_b200/csrc"
make -j8 LIBDIR=../lib_var/$1 EXTRA="$2" >/dev/null 2>&1
ls -la ../lib_var/$1/libfrequensee.so
