#!/bin/bash
# A/B of pg_tile_tc switches on VGG16 conv1_2 / conv2_2 / AllConvNet conv2 shapes: prints ms per launch (scratch/prof_conv.py)
for st in 0 1; do
  echo "KN_TILE_STAGGER=$st"
  KN_TILE_STAGGER=$st python scratch/prof_conv.py 64 64 224 256 2>&1 | tail -1
  KN_TILE_STAGGER=$st python scratch/prof_conv.py 128 128 112 256 2>&1 | tail -1
  KN_TILE_STAGGER=$st python scratch/prof_conv.py 96 96 32 4096 2>&1 | tail -1
  KN_TILE_STAGGER=$st KEYNET_B200_LIB=scratch/libkeynet_b200_prof1.so python scratch/tile_prof.py 64 64 224 256 2>&1 | head -1
done
