#!/bin/bash
# Ablations of the tensor-core candidate kernel on the config-4 workload (NRC_TC_DBG bits: 1 skip the epilogue, 8 Tensor Memory read-out only)
O=gpurun_out/${1:-r2e}; mkdir -p $O
for dbg in 0 8 1; do
  (NRC_TC_DBG=$dbg timeout 300 python bench.py --workload eval-synth --only --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('NRC_TC_DBG=$dbg: candidate kernel %.1f ms per 37 888 x 10 M x 128 launch, %.0f TFLOP/s, %.3f of the sustained bf16 peak' % (r['launch_us'] / 1e3, r['achieved'], r['frac']))") >> $O/tc_ablation.log 2>&1
done
cat $O/tc_ablation.log
