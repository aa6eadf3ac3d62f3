set -x
O=gpurun_out/$1; mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
python tests/iter_breakdown.py > $O/iter_breakdown.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-supplementary > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-supplementary > $O/ncu_launch.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
tail -2 $O/smoke.log
