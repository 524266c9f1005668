#!/bin/bash
# A/B two builds of librgie.so inside one gpurun call: tools/ab.sh "<command>"  (A = in-tree build, B = build_ab/librgie_b.so)
L=regressor_guided_image_editing_b200/librgie.so
cp $L /tmp/librgie_a.so
for round in 1 2; do
  for v in a b; do
    if [ $v = a ]; then cp /tmp/librgie_a.so $L; else cp build_ab/librgie_b.so $L; fi
    echo "== variant $v round $round"
    bash -c "$1"
  done
done
cp /tmp/librgie_a.so $L
