#!/bin/bash
# Short 1-GPU session: the sampler-queue pipeline (parity first, under a short timeout), then timings.
O=gpurun_out/${1:-r2f}; mkdir -p $O
(timeout 240 python -m pytest tests/test_gpu_epoch.py -q -m gpu -x -k "csr_fed or replicated_head" 2>&1 | tail -15) > $O/pytest_pipe.log 2>&1; echo "rc pipe $?" >> $O/rc.log
for pipe in 1 0; do
  (NRC_SGD_PIPE=$pipe timeout 400 python bench.py --only --steps 20 --warmup 5 2> $O/bench_pipe$pipe.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('NRC_SGD_PIPE=$pipe: %.3f G triplets/s, %.1f us per launch, %.3f of the HBM peak, head sync %.1f us' % (d['value'] / 1e9, r['launch_us'], r['frac'], r['replicated_head']['sync_us_mean']))") >> $O/sgd_pipe.log 2>&1
done
(NRC_SGD_PIPE=0 NRC_BENCH_N_HOT=0 timeout 400 python bench.py --only --steps 20 --warmup 5 2> /dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('register form, no replicated head: %.3f G triplets/s, %.1f us per launch, %.3f of the HBM peak' % (d['value'] / 1e9, r['launch_us'], r['frac']))") >> $O/sgd_pipe.log 2>&1
(timeout 400 python profiles/dbg_round2.py) > $O/dbg_round2.log 2>&1
(NRC_SPMM_WAVES=0 timeout 400 python profiles/dbg_round2.py 2>&1 | grep -i -E "spmm|lightgcn") > $O/dbg_spmm_waves0.log 2>&1
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40) > $O/pytest_all.log 2>&1; echo "rc pytest $?" >> $O/rc.log
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_lightgcn.csv python bench.py --workload lightgcn-gowalla --only --steps 3 --warmup 3 > $O/lgcn_under_ncu.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_sgd_pipe -s 4 -c 2 -o $O/prof_sgd_pipe python bench.py --only --steps 4 --warmup 3 > $O/ncu1.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_fast -s 4 -c 2 -o $O/prof_spmm python bench.py --workload lightgcn-gowalla --only --steps 3 --warmup 3 > $O/ncu2.log 2>&1)
cat $O/rc.log; tail -n 8 $O/pytest_pipe.log; cat $O/sgd_pipe.log $O/dbg_round2.log $O/dbg_spmm_waves0.log; tail -n 6 $O/pytest_all.log; ls $O
