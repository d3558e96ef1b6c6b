#!/bin/bash
# 1-GPU session: LightGCN region diagnostic, then smoke + the driver's two bench commands + launch list (final code).
O=gpurun_out/${1:-r2p}; mkdir -p $O
(timeout 300 python profiles/dbg_lightgcn.py) > $O/dbg_lightgcn.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; echo "rc smoke $?" >> $O/rc.log
(timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); echo "rc bench $?" >> $O/rc.log
(timeout 600 python bench.py --impl reference > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err); echo "rc bench_ref $?" >> $O/rc.log
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_headline.csv python bench.py --only --steps 4 --warmup 3 > $O/b_under_ncu.log 2>&1)
cat $O/rc.log $O/dbg_lightgcn.log; tail -2 $O/smoke.log; tail -c 200 $O/bench_n1.err
python profiles/results_table.py $O/bench_n1.json 2>/dev/null | head -8
