#!/bin/bash
# Multi-GPU session (N = $2): correctness of the row-sharded step, then the driver's bench command and the
# alternatives of its two switches (kernel form, replicated head), item-sharded evaluator, reference arm under torchrun.
O=gpurun_out/${1:-r2m}; N=${2:-2}; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29521 tests/mgpu_sharded_check.py) > $O/sharded.log 2>&1; echo "rc sharded $?" >> $O/rc.log
(timeout 600 $TR --master-port 29526 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err); echo "rc bench $?" >> $O/rc.log
(NRC_SGD_PIPE=1 timeout 600 $TR --master-port 29528 bench.py --gpus $N --only --steps 10 --warmup 3 > $O/bench_n${N}_pipe.json 2> $O/bench_n${N}_pipe.err); echo "rc bench_pipe $?" >> $O/rc.log
(NRC_BENCH_N_HOT=0 timeout 600 $TR --master-port 29527 bench.py --gpus $N --only --steps 10 --warmup 3 > $O/bench_n${N}_nohead.json 2> $O/bench_n${N}_nohead.err); echo "rc bench_nohead $?" >> $O/rc.log
(timeout 300 $TR --master-port 29531 tests/mgpu_eval_sharded_check.py) > $O/eval_sharded.log 2>&1; echo "rc eval_sharded $?" >> $O/rc.log
(timeout 600 $TR --master-port 29530 bench.py --gpus $N --workload eval-sharded --steps 4 --warmup 3 > $O/bench_eval_sharded_n$N.json 2> $O/bench_eval_sharded_n$N.err); echo "rc eval_sharded_bench $?" >> $O/rc.log
(timeout 300 $TR --master-port 29532 bench.py --impl reference --gpus $N --steps 4 --warmup 3 > $O/bench_reference_n$N.json 2> $O/bench_reference_n$N.err); echo "rc bench_ref $?" >> $O/rc.log
cat $O/rc.log; tail -n 3 $O/sharded.log; grep "item-sharded" $O/eval_sharded.log
for f in $O/bench_n${N}.json $O/bench_n${N}_pipe.json $O/bench_n${N}_nohead.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(sys.argv[1], "%.3f G triplets/s, %.3f ms/step, kernel %.0f us, head sync %.0f us, nvlink %s, e2e %.3f G/s" % (
        d["value"] / 1e9, d["ms_per_step"], r["launch_us"], r["replicated_head"]["sync_us_mean"],
        r.get("nvlink", {}).get("GBps_per_gpu_per_direction"), d["e2e"]["value"] / 1e9))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -c 300 $O/bench_n$N.err; wc -c $O/bench_reference_n$N.json
