#!/bin/bash
# A/B aid: builds libb200nav.so once per "name:flags" argument into variants/<name>.so, then restores the default build.
# On the GPU box: for v in variants/*.so; do cp $v ros_navigation_b200/csrc/libb200nav.so; python bench.py ...; done
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  make -s -B -C ros_navigation_b200/csrc EXTRA_NVFLAGS="$flags" > /dev/null
  cp ros_navigation_b200/csrc/libb200nav.so variants/$name.so
  grep -A3 "himm_tile_coded" ros_navigation_b200/csrc/ptxas.log | grep "Used\|spill" | sed "s/^/$name: /"
done
make -s -B -C ros_navigation_b200/csrc > /dev/null
