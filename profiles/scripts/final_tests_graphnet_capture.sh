O=gpurun_out/$1; mkdir -p $O
python -m pytest tests -m gpu -q > $O/tests_all.log 2>&1; tail -3 $O/tests_all.log
python bench.py --workload graphnet --steps 1 --warmup 1 --gn-epochs 1 > $O/gn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:graphnet_train_tc -s 1 -c 1 -o $O/prof_gn python bench.py --workload graphnet --steps 1 --warmup 1 --gn-epochs 1 > $O/ncu_gn.log 2>&1
ncu -i $O/prof_gn.ncu-rep --page raw --csv > $O/gn_raw.csv 2>/dev/null
ncu -i $O/prof_gn.ncu-rep --page source --csv --print-source cuda,sass > $O/gn_src.csv 2>/dev/null
rm -f $O/prof_gn.ncu-rep
ls -la $O | tail -8
