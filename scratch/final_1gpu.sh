#!/bin/bash
# final single-GPU evidence run of a round: tests, smoke, default bench, key-compile sweep, ncu launch list + full capture of the dominant launch
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final/pytest_gpu.log 2>&1; tail -2 gpurun_out/final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1; tail -1 gpurun_out/final/smoke.log
timeout 600 python bench.py > gpurun_out/final/bench_1gpu.json 2> gpurun_out/final/bench_1gpu.err; tail -c 300 gpurun_out/final/bench_1gpu.json
timeout 400 python bench.py --mode keycompile > gpurun_out/final/keycompile.json 2> gpurun_out/final/keycompile.err; tail -c 300 gpurun_out/final/keycompile.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final/vgg16_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/final/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pg_tile_tc --launch-skip 3 -c 1 -f -o gpurun_out/final/vgg16_conv1_2_tile python bench.py --no-cpu-baseline --no-extra --steps 1 --warmup 3 > gpurun_out/final/ncu_tile.log 2>&1
ls -la gpurun_out/final/*.ncu-rep gpurun_out/final/vgg16_launches.csv
