#!/bin/bash
# The round's measured artefacts, one GPU (run under gpurun; results land in gpurun_out/<tag>_*):
#   tools/final_profile.sh <tag>
# gpu tests, the default bench line, the many-particle and regional workloads, the ncu launch list and one full capture of k_fast.
tag=${1:-run}; o=gpurun_out; mkdir -p $o
timeout 900 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.txt 2>&1; tail -1 $o/${tag}_pytest.txt
python __graft_entry__.py smoke > $o/${tag}_smoke.txt 2>&1; tail -1 $o/${tag}_smoke.txt
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -c 600 $o/${tag}_bench.json
python bench.py --impl reference > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
python bench.py --workload cfg3_20w_5kp_180d_65536s --steps 5 --warmup 3 > $o/${tag}_bench_cfg3.json 2>> $o/${tag}_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $o/${tag}_plain.json 2>> $o/${tag}_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -f -o $o/${tag}_k_fast \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu2.log 2>&1
python bench.py --workload cfg4_500w_1Mp_365d_256s --steps 3 --warmup 3 > $o/${tag}_bench_cfg4.json 2>> $o/${tag}_bench.err
tail -c 400 $o/${tag}_bench_cfg4.json
ls -la $o | tail -15
