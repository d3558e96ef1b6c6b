#!/bin/bash
# Second kind of 1-GPU session of round 2: the pipelined SGD kernel (first under a short timeout), the warp-per-sample
# NCF epoch, evaluator stage timings, tensor-core ablations, launch lists.  Run under gpurun from the repo root.
O=gpurun_out/${1:-r2e}; mkdir -p $O
(timeout 180 python -m pytest tests/test_gpu_epoch.py -q -m gpu -x -k "csr_fed" 2>&1 | tail -15) > $O/pytest_pipe.log 2>&1; echo "rc pipe $?" >> $O/rc.log
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40) > $O/pytest_all.log 2>&1; echo "rc pytest $?" >> $O/rc.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > $O/smoke.log 2>&1; echo "rc smoke $?" >> $O/rc.log
(timeout 400 python profiles/dbg_round2.py) > $O/dbg_round2.log 2>&1
for pipe in 1 0; do
  (NRC_SGD_PIPE=$pipe timeout 400 python bench.py --only --steps 20 --warmup 5 2> $O/bench_pipe$pipe.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('NRC_SGD_PIPE=$pipe: %.3f G triplets/s, %.1f us per launch, %.3f of the HBM peak' % (d['value'] / 1e9, r['launch_us'], r['frac']))") >> $O/sgd_pipe.log 2>&1
done
(timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err)
bash profiles/r2_tc_ablation.sh ${1:-r2e} > /dev/null 2>&1
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_lightgcn.csv python bench.py --workload lightgcn-gowalla --only --steps 3 --warmup 3 > $O/lgcn_under_ncu.log 2>&1)
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_neumf_eval.csv env NRC_EVAL_ONLY=1 python profiles/dbg_round2.py > $O/neumf_under_ncu.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_sgd_pipe -s 4 -c 2 -o $O/prof_sgd_pipe python bench.py --only --steps 4 --warmup 3 > $O/ncu1.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:ncf_epoch -s 1 -c 1 -o $O/prof_ncf_epoch python bench.py --workload neumf-ml100k --only --steps 400 --warmup 3 > $O/ncu4.log 2>&1)
cat $O/rc.log; tail -n 8 $O/pytest_pipe.log; tail -n 6 $O/pytest_all.log; tail -n 3 $O/smoke.log; cat $O/dbg_round2.log $O/sgd_pipe.log $O/tc_ablation.log; tail -c 300 $O/bench_n1.err; ls $O
