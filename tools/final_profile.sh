#!/bin/bash
# The round's measured artefacts, one GPU (run under gpurun; results land in gpurun_out/<tag>_*):
#   tools/final_profile.sh <tag>
# gpu tests, smoke, the default bench line (config 2 + the config-3 block), the reference arm, configs 3 / 4 / 5 as their own
# lines, the ncu launch list of the default command and one full capture of the bit-sliced kernel on configs 2, 3, 4 and 5.
tag=${1:-run}; o=gpurun_out; mkdir -p $o
timeout 1200 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.txt 2>&1; tail -1 $o/${tag}_pytest.txt
python __graft_entry__.py smoke > $o/${tag}_smoke.txt 2>&1; tail -1 $o/${tag}_smoke.txt
python bench.py > $o/${tag}_bench.json 2> $o/${tag}_bench.err; tail -c 400 $o/${tag}_bench.json
python bench.py --impl reference > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
python bench.py --workload cfg3_20w_5kp_180d_65536s --steps 5 --warmup 3 > $o/${tag}_bench_cfg3.json 2>> $o/${tag}_bench.err
python bench.py --workload cfg4_500w_1Mp_365d_256s --steps 3 --warmup 3 > $o/${tag}_bench_cfg4.json 2>> $o/${tag}_bench.err
python bench.py --workload cfg5_5kw_10Mp_730d_32s --steps 3 --warmup 2 --no-cpu-baseline > $o/${tag}_bench_cfg5.json 2>> $o/${tag}_bench.err
python tools/trajectory_time.py 64 > $o/${tag}_trajectory.txt 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-dropin"
$B > $o/${tag}_plain.json 2>> $o/${tag}_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv $B > $o/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_bits -s 3 -c 1 -f -o $o/${tag}_k_bits_cfg2 $B > $o/${tag}_ncu2.log 2>&1
B3="python bench.py --workload cfg3_20w_5kp_180d_65536s --steps 2 --warmup 1 --no-cpu-baseline"
$B3 > $o/${tag}_plain3.json 2>> $o/${tag}_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:k_bits -s 1 -c 1 -f -o $o/${tag}_k_bits_cfg3 $B3 > $o/${tag}_ncu3.log 2>&1
B4="python bench.py --workload cfg4_500w_1Mp_365d_256s --steps 2 --warmup 1 --no-cpu-baseline"
$B4 > $o/${tag}_plain4.json 2>> $o/${tag}_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:k_bits -s 1 -c 1 -f -o $o/${tag}_k_bits_cfg4 $B4 > $o/${tag}_ncu4.log 2>&1
B5="python bench.py --workload cfg5_5kw_10Mp_730d_32s --steps 2 --warmup 1 --no-cpu-baseline"
$B5 > $o/${tag}_plain5.json 2>> $o/${tag}_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:k_bits -s 1 -c 1 -f -o $o/${tag}_k_bits_cfg5 $B5 > $o/${tag}_ncu5.log 2>&1
# summaries on the box; the reports themselves (12 MB each) stay behind except config 2's: gpurun_out/ comes back only below 64 MiB
for c in "cfg2 3.65e9 cfg2_30w_10kp_365d_1000s" "cfg3 5.8982e10 cfg3_20w_5kp_180d_65536s" "cfg4 9.344e10 cfg4_500w_1Mp_365d_256s" "cfg5 2.336e11 cfg5_5kw_10Mp_730d_32s"; do
  set -- $c; rep=$o/${tag}_k_bits_$1.ncu-rep
  [ -f $rep ] || continue
  python tools/ncu_summary.py $rep $2 > $o/${tag}_k_bits_$1_summary.txt 2>/dev/null
  ncu -i $rep --page details > $o/${tag}_k_bits_$1_details.txt 2>/dev/null
  [ "$1" = cfg2 ] || rm -f $rep
done
ls -la $o | tail -30
