#!/bin/bash
# N-GPU session after the bulk-reduce local updates: the sharded correctness stages, the driver's bench command, and
# the same with vector REDs on local rows.
O=gpurun_out/${1:-r2o}; N=${2:-2}; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29521 tests/mgpu_sharded_check.py) > $O/sharded.log 2>&1; echo "rc sharded $?" >> $O/rc.log
(timeout 400 $TR --master-port 29526 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err); echo "rc bench $?" >> $O/rc.log
(NRC_SGD_LOCAL_BULK=0 timeout 400 $TR --master-port 29528 bench.py --gpus $N --only --steps 20 --warmup 5 > $O/bench_n${N}_red.json 2> $O/bench_n${N}_red.err); echo "rc bench_red $?" >> $O/rc.log
cat $O/rc.log; tail -n 3 $O/sharded.log
for f in $O/bench_n${N}.json $O/bench_n${N}_red.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(sys.argv[1], "%.3f G triplets/s, %.3f ms/step, kernel %.0f us, head sync %.0f us, nvlink %s, e2e %.3f G/s, warm-up %d" % (
        d["value"] / 1e9, d["ms_per_step"], r["launch_us"], r["replicated_head"]["sync_us_mean"],
        r.get("nvlink", {}).get("GBps_per_gpu_per_direction"), d["e2e"]["value"] / 1e9, d["warmup"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -c 300 $O/bench_n$N.err
