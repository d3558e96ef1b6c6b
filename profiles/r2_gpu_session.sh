#!/bin/bash
# One GPU session of round 2: parity suite, kernel timings, bench, ncu evidence.  Run under gpurun from the repo root.
O=gpurun_out/${1:-r2c}; mkdir -p $O
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40) > $O/pytest_all.log 2>&1
(timeout 300 python profiles/dbg_epoch.py) > $O/dbg_epoch.log 2>&1
(timeout 400 python profiles/dbg_round2.py) > $O/dbg_round2.log 2>&1
(NRC_SPMM_UN=8 timeout 400 python profiles/dbg_round2.py 2>&1 | grep -i -E "spmm|lightgcn") > $O/dbg_spmm_un8.log 2>&1
for ch in 1 2; do
  (NRC_TC_CH=$ch timeout 300 python bench.py --workload eval-synth --only --steps 3 --warmup 3 > $O/eval_synth_ch$ch.json 2> $O/eval_synth_ch$ch.err)
done
(timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err)
(timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err)
# the remote-row update flavours of the sharded kernel, exercised on LOCAL memory (NRC_FORCE_REMOTE_PATH): parity + speed
for m in 0 1 2; do
  (NRC_FORCE_REMOTE_PATH=1 NRC_PEER_VEC_RED=$m timeout 300 python -m pytest tests/test_gpu_epoch.py -q -m gpu -k "csr_fed" 2>&1 | tail -3) > $O/remote_mode$m.log 2>&1
  (NRC_FORCE_REMOTE_PATH=1 NRC_PEER_VEC_RED=$m timeout 300 python bench.py --only --steps 10 --warmup 3 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mode $m', d['value'], d['roofline']['frac'])") >> $O/remote_mode$m.log 2>&1
done
# ncu: launch list of the headline command, then full captures of the dominant kernels
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_headline.csv python bench.py --only --steps 4 --warmup 3 > $O/b_under_ncu.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_sgd_stream -s 4 -c 2 -o $O/prof_sgd_stream python bench.py --only --steps 4 --warmup 3 > $O/ncu1.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_fast -s 4 -c 2 -o $O/prof_spmm python bench.py --workload lightgcn-gowalla --only --steps 3 --warmup 3 > $O/ncu2.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_epoch_kernel -s 1 -c 1 -o $O/prof_mf_epoch python bench.py --workload bprmf-ml100k --only --steps 157 --warmup 3 > $O/ncu3.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:ncf_epoch -s 1 -c 1 -o $O/prof_ncf_epoch python bench.py --workload neumf-ml100k --only --steps 400 --warmup 3 > $O/ncu4.log 2>&1)
tail -n 6 $O/pytest_all.log; cat $O/dbg_epoch.log $O/dbg_round2.log $O/remote_mode*.log; tail -c 300 $O/bench_n1.err; ls -la $O
