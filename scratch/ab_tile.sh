#!/bin/bash
# A/B of a pg_tile_tc switch ($1 = environment variable name) on VGG16 conv1_2 / conv2_1 / conv2_2 / AllConvNet conv2 shapes (scratch/prof_conv.py)
for v in 0 1; do
  echo "$1=$v"
  env $1=$v python scratch/prof_conv.py 64 64 224 256 2>&1 | tail -1
  env $1=$v python scratch/prof_conv.py 64 128 112 256 2>&1 | tail -1
  env $1=$v python scratch/prof_conv.py 128 128 112 256 2>&1 | tail -1
  env $1=$v python scratch/prof_conv.py 96 96 32 4096 2>&1 | tail -1
done
