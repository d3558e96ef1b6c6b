#!/bin/bash
# Multi-GPU session (N = $2): hot-row probe, sharded parity with the replicated head, weak-scaling bench with and
# without the head / the pipelined kernel, item-sharded evaluator.
O=gpurun_out/${1:-r2q}; N=${2:-2}; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 profiles/peer_probe.bin 12500000 2097152 1) > $O/peer_probe.log 2>&1; echo "rc probe $?" >> $O/rc.log
(timeout 300 $TR --master-port 29521 tests/mgpu_sharded_check.py ipc) > $O/sharded_ipc.log 2>&1; echo "rc ipc $?" >> $O/rc.log
(NRC_SGD_PIPE=0 timeout 300 $TR --master-port 29522 tests/mgpu_sharded_check.py ipc) > $O/sharded_ipc_register.log 2>&1; echo "rc ipc_register $?" >> $O/rc.log
(timeout 600 $TR --master-port 29526 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err); echo "rc bench $?" >> $O/rc.log
(NRC_BENCH_N_HOT=0 timeout 600 $TR --master-port 29527 bench.py --gpus $N --only --steps 10 --warmup 3 > $O/bench_n${N}_nohead.json 2> $O/bench_n${N}_nohead.err); echo "rc bench_nohead $?" >> $O/rc.log
(NRC_SGD_PIPE=0 timeout 600 $TR --master-port 29528 bench.py --gpus $N --only --steps 10 --warmup 3 > $O/bench_n${N}_register.json 2> $O/bench_n${N}_register.err); echo "rc bench_register $?" >> $O/rc.log
(NRC_BENCH_N_HOT=262144 timeout 600 $TR --master-port 29529 bench.py --gpus $N --only --steps 10 --warmup 3 > $O/bench_n${N}_head256k.json 2> $O/bench_n${N}_head256k.err); echo "rc bench_head256k $?" >> $O/rc.log
(timeout 600 $TR --master-port 29530 bench.py --gpus $N --workload eval-sharded --steps 4 --warmup 3 > $O/bench_eval_sharded_n$N.json 2> $O/bench_eval_sharded_n$N.err); echo "rc eval_sharded_bench $?" >> $O/rc.log
cat $O/rc.log; cat $O/peer_probe.log; tail -n 8 $O/sharded_ipc.log; tail -n 3 $O/sharded_ipc_register.log
for f in $O/bench_n${N}.json $O/bench_n${N}_nohead.json $O/bench_n${N}_register.json $O/bench_n${N}_head256k.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
    print(sys.argv[1], "%.3f G triplets/s, %.3f ms/step, kernel %.0f us, head sync %.0f us, nvlink %s" % (
        d["value"] / 1e9, d["ms_per_step"], r["launch_us"], r["replicated_head"]["sync_us_mean"], r.get("nvlink", {}).get("GBps_per_gpu_per_direction")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
tail -c 400 $O/bench_n$N.err
