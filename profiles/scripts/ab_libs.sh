# A/B timing of alternative builds of libddrl_b200.so in ONE gpurun call (same box, same clocks): every variant is a shared library
# under ab/ (git-ignored), selected through DDRL_B200_LIB; optional environment switches after a second colon.
# usage: bash profiles/scripts/ab_libs.sh <outdir under gpurun_out> name:lib[:ENV=1] ...
mkdir -p gpurun_out/$1
shift_out=$1; shift
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; lib=${rest%%:*}; envs=${rest#*:}
  ( export DDRL_B200_LIB=$PWD/ab/lib_$lib.so; [ "$envs" != "$lib" ] && [ -n "$envs" ] && export $envs
    python tests/phase_clock.py > gpurun_out/$shift_out/phase_$name.txt 2>&1
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-supplementary > gpurun_out/$shift_out/bench_$name.json 2> gpurun_out/$shift_out/bench_$name.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/$shift_out/bench_$name.json').read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], d['roofline']['us_per_sgd_step'], d.get('ranks_identical'))"
    grep -h "mean phase" gpurun_out/$shift_out/phase_$name.txt )
done
