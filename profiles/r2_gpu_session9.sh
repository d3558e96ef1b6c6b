#!/bin/bash
# 1-GPU session: local rows updated by bulk reduce-adds (NRC_SGD_LOCAL_BULK, default 1) vs vector REDs: parity, speed,
# ncu capture; then the whole GPU suite and the driver's bench command.
O=gpurun_out/${1:-r2l}; mkdir -p $O
(timeout 400 python -m pytest tests/test_gpu_epoch.py -q -m gpu -k "csr_fed or replicated_head or lazy" 2>&1 | tail -15) > $O/pytest_sel.log 2>&1
for b in 1 0; do
  (NRC_SGD_LOCAL_BULK=$b timeout 300 python bench.py --only --steps 20 --warmup 5 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('NRC_SGD_LOCAL_BULK=$b: %.3f G triplets/s, %.1f us per launch (min %.1f), %.3f of the HBM peak, e2e %.3f G/s; lazy Adam %.3f G/s %.3f' % (d['value']/1e9, r['launch_us'], r['launch_us_min'], r['frac'], d['e2e']['value']/1e9, d['lazy_adam']['value']/1e9, d['lazy_adam']['roofline']['frac']))") >> $O/local_bulk.log 2>&1
done
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_sgd_stream -s 4 -c 1 -o $O/prof_sgd_stream_bulk python bench.py --only --steps 4 --warmup 3 > $O/ncu1.log 2>&1)
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12) > $O/pytest_all.log 2>&1
(timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); echo "rc bench $?" >> $O/rc.log
cat $O/rc.log; tail -n 6 $O/pytest_sel.log; cat $O/local_bulk.log; tail -n 5 $O/pytest_all.log; tail -c 200 $O/bench_n1.err
python profiles/results_table.py $O/bench_n1.json 2>/dev/null | head -8
