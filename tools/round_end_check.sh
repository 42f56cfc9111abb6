#!/bin/bash
# One gpurun call that re-verifies HEAD on a fresh B200: GPU tests, smoke, default bench, the strong-scaling shard shape,
# the launch list and one `ncu --set full` capture of the dominant kernel (each ncu pass after the plain run exited 0).
#   tools/round_end_check.sh <tag> [tests|notests]
tag=${1:-chk}
o=gpurun_out
mkdir -p $o
if [ "${2:-tests}" = tests ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > $o/${tag}_gputest.log 2>&1; echo "pytest rc=$?"
  tail -3 $o/${tag}_gputest.log
  timeout 90 python __graft_entry__.py --smoke > $o/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
fi
timeout 300 python bench.py > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench rc=$?"
cut -c1-400 $o/${tag}_bench_n1.json
timeout 120 python bench.py --cells 12500 --steps 20 --warmup 5 --no-cpu --no-e2e > $o/${tag}_bench_12k.json 2>/dev/null
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${tag}_plain.log 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${tag}_ncu_l.log 2>&1
for k in ${NCU_KERNELS:-k_sweep}; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 4 --launch-count 1 -f \
      -o $o/${tag}_prof_$k python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $o/${tag}_ncu_$k.log 2>&1
done
echo done
