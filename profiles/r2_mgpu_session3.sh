#!/bin/bash
# 2-GPU session: how fast is the product's peer mapping (IPC / symmetric memory) on uniform ids, then the bench again
O=gpurun_out/${1:-r2r}; N=${2:-2}; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29540 tests/mgpu_peer_rate.py ipc) > $O/peer_rate_ipc.log 2>&1; echo "rc rate_ipc $?" >> $O/rc.log
(timeout 300 $TR --master-port 29541 tests/mgpu_peer_rate.py symm) > $O/peer_rate_symm.log 2>&1; echo "rc rate_symm $?" >> $O/rc.log
(timeout 600 $TR --master-port 29526 bench.py --gpus $N --only --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err); echo "rc bench $?" >> $O/rc.log
cat $O/rc.log; grep "item rows" $O/peer_rate_ipc.log $O/peer_rate_symm.log; tail -n 5 $O/peer_rate_symm.log
python - "$O/bench_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(sys.argv[1], "%.3f G triplets/s, %.3f ms/step, kernel %.0f us, head sync %.0f us" % (
        d["value"] / 1e9, d["ms_per_step"], r["launch_us"], r["replicated_head"]["sync_us_mean"]))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
