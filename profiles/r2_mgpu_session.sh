#!/bin/bash
# Multi-GPU session of round 2 (N = $2, default 2): peer-memory checks, item-sharded evaluator, weak-scaling bench.
O=gpurun_out/${1:-r2m}; N=${2:-2}; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29521 tests/mgpu_sharded_check.py ipc) > $O/sharded_ipc.log 2>&1; echo "rc ipc $?" >> $O/rc.log
(timeout 300 $TR --master-port 29522 tests/mgpu_sharded_check.py symm) > $O/sharded_symm.log 2>&1; echo "rc symm $?" >> $O/rc.log
(NRC_PEER_VEC_RED=1 timeout 300 $TR --master-port 29523 tests/mgpu_sharded_check.py ipc) > $O/sharded_ipc_vecred.log 2>&1; echo "rc ipc_vecred $?" >> $O/rc.log
(NRC_PEER_VEC_RED=2 timeout 300 $TR --master-port 29533 tests/mgpu_sharded_check.py ipc) > $O/sharded_ipc_bulkred.log 2>&1; echo "rc ipc_bulkred $?" >> $O/rc.log
(timeout 300 $TR --master-port 29524 tests/mgpu_eval_sharded_check.py) > $O/eval_sharded.log 2>&1; echo "rc eval_sharded $?" >> $O/rc.log
(timeout 300 $TR --master-port 29525 tests/mgpu_eval_check.py) > $O/eval_users.log 2>&1; echo "rc eval_users $?" >> $O/rc.log
(timeout 600 $TR --master-port 29526 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err); echo "rc bench $?" >> $O/rc.log
(NRC_PEER_VEC_RED=1 timeout 600 $TR --master-port 29527 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n${N}_vecred.json 2> $O/bench_n${N}_vecred.err); echo "rc bench_vecred $?" >> $O/rc.log
(NRC_PEER_VEC_RED=2 timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n${N}_bulkred.json 2> $O/bench_n${N}_bulkred.err); echo "rc bench_bulkred $?" >> $O/rc.log
(NRC_PEER_BACKEND=symm timeout 600 $TR --master-port 29528 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n${N}_symm.json 2> $O/bench_n${N}_symm.err); echo "rc bench_symm $?" >> $O/rc.log
(timeout 600 $TR --master-port 29529 bench.py --gpus $N --workload eval-sharded --steps 4 --warmup 3 > $O/bench_eval_sharded_n$N.json 2> $O/bench_eval_sharded_n$N.err); echo "rc eval_sharded_bench $?" >> $O/rc.log
(timeout 600 $TR --master-port 29530 bench.py --gpus $N --workload eval-synth --steps 4 --warmup 3 > $O/bench_eval_synth_n$N.json 2> $O/bench_eval_synth_n$N.err); echo "rc eval_synth_bench $?" >> $O/rc.log
cat $O/rc.log; tail -n 8 $O/sharded_ipc.log $O/sharded_symm.log $O/sharded_ipc_vecred.log $O/sharded_ipc_bulkred.log $O/eval_sharded.log; head -c 700 $O/bench_n$N.json; tail -c 600 $O/bench_n$N.err
