#!/bin/bash
# 2-GPU probe session: what limits random 512-byte rows over NVLink, + the item-sharded evaluator check
O=gpurun_out/${1:-r2p}; mkdir -p $O
nvidia-smi topo -m > $O/topo.log 2>&1
(timeout 300 profiles/peer_probe.bin) > $O/peer_probe.log 2>&1; echo "rc probe $?" >> $O/rc.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
(timeout 300 $TR --master-port 29524 tests/mgpu_eval_sharded_check.py) > $O/eval_sharded.log 2>&1; echo "rc eval_sharded $?" >> $O/rc.log
cat $O/rc.log; cat $O/peer_probe.log; grep "item-sharded" $O/eval_sharded.log; head -20 $O/topo.log
