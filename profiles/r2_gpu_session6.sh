#!/bin/bash
# 1-GPU session: the 8(f) rank 3-4 kernels (new tests), then the driver's bench command with the final defaults.
O=gpurun_out/${1:-r2i}; mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_extras.py -q -m gpu 2>&1 | tail -60) > $O/pytest_extras.log 2>&1
(timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); echo "rc bench $?" >> $O/rc.log
cat $O/rc.log; tail -n 60 $O/pytest_extras.log; tail -c 300 $O/bench_n1.err
python - $O/bench_n1.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
print("%.3f G triplets/s, %.1f us per launch, %.3f of the HBM peak (%s), e2e %.3f G/s, launches %d" % (
    d["value"] / 1e9, r["launch_us"], r["frac"], r["kernel"], d["e2e"]["value"] / 1e9, d["gpu_launches"]))
PY
