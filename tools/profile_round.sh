#!/bin/bash
# One GPU-box pass that produces everything profiles/ is built from (run under gpurun from the repo root):
#   1. the plain bench lines (never under a profiler),
#   2. per-call CUDA-event profiles of one train step / one sampling pass (tools/prof_layers.py),
#   3. ncu launch lists (gpu__time_duration.sum) of two Generator sampling passes and of two train steps,
#   4. `ncu --set full` captures of the four heaviest layers of the sampling pass, each launched on its own
#      (tools/prof_kernel.py: bench.py's ROOFLINE_LAYERS).
set -u
R=${1:-r02}
O=gpurun_out
python bench.py > $O/bench_${R}.json 2> $O/bench_${R}.err || { tail -5 $O/bench_${R}.err; exit 1; }
python bench.py --workload train --steps 20 --no-extras > $O/bench_${R}_train.json 2>> $O/bench_${R}.err
python bench.py --workload train --steps 20 --no-extras --no-graph > $O/bench_${R}_train_eager.json 2>> $O/bench_${R}.err
python bench.py --workload train --events 1 --steps 20 --no-extras > $O/bench_${R}_train1.json 2>> $O/bench_${R}.err
python bench.py --workload attn-sweep > $O/bench_${R}_sweep.json 2>> $O/bench_${R}.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${R}_ref.json 2>> $O/bench_${R}.err
python tools/prof_layers.py train 8 400 > $O/layers_train_full.txt 2>/dev/null
python tools/prof_layers.py sample 16 200 > $O/layers_sample.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_sample.csv \
    python tools/prof_sample.py 16 2 > $O/ncu_sample.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $O/launches_train.csv \
    python tools/prof_train.py 8 2 > $O/ncu_train.log 2>&1
# the four heaviest layers of the sampling pass, each launched alone (third launch = warm): DRAM traffic, stalls
for T in l32_64 l32_1 l16_32 l16_16; do
  ncu --set full --clock-control none --import-source on -k regex:conv_t -s 2 -c 1 -f -o $O/prof_$T \
      python tools/prof_kernel.py $T 3 > $O/ncu_$T.log 2>&1
done
# summaries are made HERE (the .ncu-rep files together exceed what gpurun copies back) and travel as text
python tools/make_profiles.py $R > $O/make_profiles.log 2>&1
python tools/profiles_readme.py $R >> $O/make_profiles.log 2>&1
rm -rf $O/profiles_new && cp -r profiles $O/profiles_new
rm -f $O/*.ncu-rep
ls -la $O | tail -12
