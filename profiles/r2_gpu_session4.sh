#!/bin/bash
# Short 1-GPU session: pipeline v3 (bulk-copy loads + register REDs) parity and speed, NCF predict kernel profile
O=gpurun_out/${1:-r2g}; mkdir -p $O
(timeout 240 python -m pytest tests/test_gpu_epoch.py tests/test_gpu_ncf.py -q -m gpu -x -k "csr_fed or replicated_head or ncf_scores" 2>&1 | tail -15) > $O/pytest_pipe.log 2>&1; echo "rc pipe $?" >> $O/rc.log
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
(NRC_EVAL_ONLY=1 timeout 300 python profiles/dbg_round2.py) > $O/dbg_eval.log 2>&1
(timeout 400 python profiles/dbg_round2.py 2>&1 | grep -i -E "spmm|lightgcn") > $O/dbg_spmm.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:ncf_scores_tile -c 1 -o $O/prof_ncf_scores env NRC_EVAL_ONLY=1 python profiles/dbg_round2.py > $O/ncu5.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_sgd_pipe -s 4 -c 1 -o $O/prof_sgd_pipe python bench.py --only --steps 4 --warmup 3 > $O/ncu1.log 2>&1)
cat $O/rc.log; tail -n 6 $O/pytest_pipe.log; cat $O/sgd_pipe.log $O/dbg_eval.log $O/dbg_spmm.log; ls $O
