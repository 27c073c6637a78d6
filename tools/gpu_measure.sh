#!/bin/bash
# Runs ON the GPU box (under gpurun): GPU tests, the bench line, the ncu launch list of the SAME bench command and one
# ncu --set full capture of the two rollout kernels at bench size.  Usage: tools/gpu_measure.sh <tag>   (e.g. r1)
# Outputs land in gpurun_out/; tools/summarise_profiles.py turns them into profiles/<tag>_*.
tag=${1:-r1}
out=gpurun_out
python -m pytest tests -m gpu -q > $out/${tag}_gpu_tests.log 2>&1; tail -1 $out/${tag}_gpu_tests.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
python bench.py --impl reference --steps 4 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
# launch list of the same command (a second plain run directly before ncu, as the profiling recipe asks)
python bench.py --no-cpu-baseline > $out/${tag}_plain_bench.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --no-cpu-baseline > $out/${tag}_ncu_launches.log 2>&1
# one launch of each: the rollout kernel (24 steps), then the per-call pair (tools/prof_rollout.py runs plan 6, then 4)
python tools/prof_rollout.py 125000 1512 24 > $out/${tag}_plain_prof.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"rollout_loop|ctrl_step|physics_step" -s 63 -c 3 -f -o $out/${tag}_prof \
      python tools/prof_rollout.py 125000 1512 24 > $out/${tag}_ncu_prof.log 2>&1
tail -2 $out/${tag}_ncu_prof.log
