#!/bin/bash
# One GPU-box pass that produces everything profiles/ is built from (run under gpurun from the repo root):
#   1. the plain bench line (never under a profiler),
#   2. ncu launch lists (gpu__time_duration.sum) of two Generator sampling passes and of two train steps,
#   3. one `ncu --set full` capture of the dominant kernel (macro-tile tcgen05 conv, 16->16 3x3 @256^2).
set -u
O=gpurun_out
python bench.py > $O/bench_r01.json 2> $O/bench_r01.err || { tail -5 $O/bench_r01.err; exit 1; }
python bench.py --workload train --steps 5 --warmup 3 --no-extras > $O/bench_r01_train.json 2>> $O/bench_r01.err
python bench.py --hbase 3 --events 8 --steps 5 --no-extras > $O/bench_r01_hbase3.json 2>> $O/bench_r01.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>> $O/bench_r01.err
python tools/prof_layers.py train 8 400 > $O/layers_train_full.txt 2>/dev/null
python tools/prof_layers.py sample 16 200 > $O/layers_sample.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_sample.csv \
    python tools/prof_sample.py 16 2 > $O/ncu_sample.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_train.csv \
    python tools/prof_train.py 8 2 > $O/ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_thin_kernel -s 12 -c 12 -f -o $O/prof_thin \
    python tools/prof_sample.py 4 2 > $O/ncu_full.log 2>&1
ls -la $O | tail -12
