#!/bin/bash
# Final measurement set of a round on ONE B200 (run under gpurun from the repo root): GPU tests, the default bench line, the
# launch list and one `ncu --set full` pass over the hot kernels of one evaluation.  Everything lands in gpurun_out/<tag>_*.
#   tools/final_measure.sh <tag> [tests|bench|launches|full|multipart|fp32|ref ...]     (default: all but ref)
TAG=${1:-r02c}; shift
WHAT=${@:-tests bench launches full multipart fp32}
O=gpurun_out
for w in $WHAT; do case $w in
tests)     timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/${TAG}_gputest.log; cat $O/${TAG}_gputest.log ;;
bench)     timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench.json ;;
multipart) timeout 600 python bench.py --workload multipart --events 2048 --no-extra --no-cpu-baseline --steps 3 > $O/${TAG}_multipart.json 2>> $O/${TAG}_bench.err ;;
fp32)      timeout 600 python bench.py --precision fp32 --no-extra --no-cpu-baseline --steps 2 > $O/${TAG}_fp32.json 2>> $O/${TAG}_bench.err ;;
ref)       timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_ref.json 2>> $O/${TAG}_bench.err ;;
launches)  CMD="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline"
           $CMD > $O/${TAG}_short.json 2>> $O/${TAG}_bench.err && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1 ;;
full)      CMD="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu-baseline"
           # one evaluation's worth of the hot kernels: skip the launches of the first evaluations (graph capture / warm-up), then 20 launches
           $CMD > /dev/null 2>> $O/${TAG}_bench.err && timeout 1200 ncu --set full --clock-control none -k regex:'embed_tc|context_rows|gemm_f32_big|modpq|layer_chain|attn3|head_' -s 126 -c 19 -f -o $O/${TAG}_full $CMD > $O/${TAG}_ncu_full.log 2>&1
           # the report itself (source view of 18 launches) is larger than what gpurun copies back: keep the per-launch counter summary
           python tools/ncu_summary.py $O/${TAG}_full.ncu-rep > $O/${TAG}_ncu_kernels.csv 2>> $O/${TAG}_bench.err; rm -f $O/${TAG}_full.ncu-rep
           tail -2 $O/${TAG}_ncu_full.log; wc -l $O/${TAG}_ncu_kernels.csv ;;
esac; done
