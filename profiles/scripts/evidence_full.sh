set -x
O=gpurun_out/$1; mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
python tests/phase_clock.py > $O/phase_clock.txt 2>&1
python tests/grad_error_table.py > $O/grad_error_table.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-supplementary > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-supplementary > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fcnet_train_tc2_kernel -s 2 -c 1 -o $O/prof_tc2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-supplementary > $O/ncu_full.log 2>&1
ncu -i $O/prof_tc2.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/prof_tc2.ncu-rep --page source --csv --print-source cuda,sass > $O/src.csv 2>/dev/null
ls -la $O
