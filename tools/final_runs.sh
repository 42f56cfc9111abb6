#!/bin/bash
# Final single-GPU records of a round (run under gpurun from the repo root).
#   tools/final_runs.sh ncu   <tag>   launch list + one `ncu --set full` capture per hot kernel (each only after the same
#                                     command has exited 0 without ncu)
#   tools/final_runs.sh bench <tag>   default bench, non-headline shapes, do_impute / C1 / C2 timings
# Outputs land in gpurun_out/<tag>_*.
what=${1:-bench}
tag=${2:-final}
o=gpurun_out
if [ "$what" = ncu ]; then
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${tag}_plain.log 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${tag}_ncu_l.log 2>&1
  for k in k_sweep k_proj k_cell_fast k_mstats; do
    ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 4 --launch-count 1 -f -o $o/${tag}_prof_$k \
        python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${tag}_ncu_$k.log 2>&1
  done
else
  python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err || exit 1
  python bench.py --cells 12500 --steps 20 --warmup 5 --no-cpu --no-e2e > $o/${tag}_bench_12k.json 2>/dev/null
  python bench.py --cells 125000 --genes 5000 --K0 20 --M0 20 --steps 5 --warmup 3 --no-cpu --no-e2e > $o/${tag}_bench_c4shard.json 2>/dev/null
  python bench.py --cells 200000 --K0 30 --M0 20 --steps 5 --warmup 3 --no-cpu --no-e2e > $o/${tag}_bench_c5max.json 2>/dev/null
  python tools/bench_mi.py > $o/${tag}_mi.json 2> $o/${tag}_mi.err
  python tools/bench_c1_c2.py > $o/${tag}_c1_c2.json 2> $o/${tag}_c1_c2.err
fi
echo done
