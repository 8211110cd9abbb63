#!/bin/bash
# A/B of a pg_tile_tc switch: scratch/ab_tile.sh VAR v1 v2 ... on VGG16 conv1_2 / conv2_1 / conv2_2 / AllConvNet conv2 shapes (scratch/prof_conv.py)
var=$1; shift
for v in "$@"; do
  echo "$var=$v"
  env $var=$v python scratch/prof_conv.py 64 64 224 256 2>&1 | tail -1
  env $var=$v python scratch/prof_conv.py 64 128 112 256 2>&1 | tail -1
  env $var=$v python scratch/prof_conv.py 128 128 112 256 2>&1 | tail -1
  env $var=$v python scratch/prof_conv.py 96 96 32 4096 2>&1 | tail -1
done
