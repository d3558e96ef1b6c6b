#!/bin/bash
# 1-GPU session: lazy-Adam kernel with nine loads in flight, SpectralCF GEMM v2, e2e with the smaller upload, the
# bulk-reduce update flavour on local rows, then the driver's bench command with the final code.
O=gpurun_out/${1:-r2k}; mkdir -p $O
(timeout 400 python -m pytest tests/test_gpu_epoch.py tests/test_gpu_extras.py -q -m gpu -k "lazy or csr_fed or spectral or row_ids" 2>&1 | tail -15) > $O/pytest_sel.log 2>&1
(timeout 300 python profiles/dbg_spectral.py) > $O/dbg_spectral.log 2>&1
for m in 1 2; do
  (NRC_FORCE_REMOTE_PATH=1 NRC_PEER_VEC_RED=$m timeout 300 python bench.py --only --steps 20 --warmup 5 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('item rows through the remote-update path on local memory, mode $m (1 vector RED, 2 bulk reduce-add):', d['value'], d['roofline']['frac'], d['roofline']['launch_us'])") >> $O/remote_modes.log 2>&1
done
(timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); echo "rc bench $?" >> $O/rc.log
cat $O/rc.log; tail -n 8 $O/pytest_sel.log; cat $O/dbg_spectral.log $O/remote_modes.log; tail -c 300 $O/bench_n1.err
python profiles/results_table.py $O/bench_n1.json 2>/dev/null | head -12
