#!/bin/bash
# usage (GPU box, via gpurun): profiles/run_final.sh <tag>
# The round's record: bench lines of configs 2 / 3 / 4 and the reference arm, the ncu launch list of the bench command and
# one `ncu --set full` capture of the dominant kernel (k_fused).  Everything lands in gpurun_out/<tag>_*.
tag=$1
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py --config 3 --no-cpu-baseline > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err
python bench.py --config 4 --no-cpu-baseline > gpurun_out/${tag}_bench_config4.json 2> gpurun_out/${tag}_bench_config4.err
python bench.py --ragged --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_bench_ragged.json 2> gpurun_out/${tag}_bench_ragged.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${tag}_launches.log 2>&1
profiles/run_ncu.sh k_fused ${tag}_fused
for f in bench bench_reference bench_config3 bench_config4 bench_ragged; do head -c 400 gpurun_out/${tag}_$f.json; echo; done
