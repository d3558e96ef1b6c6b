#!/bin/bash
# Round-2 evidence session (1 GPU), most important first: parity suite, smoke, the driver's two bench commands,
# pipelined vs register form of the CSR-fed SGD kernel, ncu launch list + full captures of the dominant kernels.
O=gpurun_out/${1:-r2h}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.log 2>&1
(timeout 240 python -m pytest tests/test_gpu_epoch.py -q -m gpu -x -k "csr_fed or replicated_head" 2>&1 | tail -15) > $O/pytest_pipe.log 2>&1
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40) > $O/pytest_all.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; echo "rc smoke $?" >> $O/rc.log
(timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err); echo "rc bench $?" >> $O/rc.log
(timeout 600 python bench.py --impl reference > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err); echo "rc bench_ref $?" >> $O/rc.log
(NRC_SGD_PIPE=0 timeout 400 python bench.py --only --steps 20 --warmup 5 2> $O/bench_pipe0.err > $O/bench_pipe0.json)
(NRC_SGD_PIPE=0 NRC_BENCH_N_HOT=0 timeout 400 python bench.py --only --steps 20 --warmup 5 2> /dev/null > $O/bench_pipe0_nohead.json)
(NRC_BENCH_N_HOT=0 timeout 400 python bench.py --only --steps 20 --warmup 5 2> /dev/null > $O/bench_pipe1_nohead.json)
python - $O <<'PY' > $O/sgd_forms.log 2>&1
import json, sys, os
O = sys.argv[1]
for name, f in (("default (pipe, head)", "bench_n1.json"), ("register form, head", "bench_pipe0.json"),
                ("register form, no head", "bench_pipe0_nohead.json"), ("pipe, no head", "bench_pipe1_nohead.json")):
    try:
        d = json.loads(open(os.path.join(O, f)).read().strip().splitlines()[-1]); r = d["roofline"]
        print("%-24s %.3f G triplets/s, %.1f us per launch, %.3f of the HBM peak, head sync %.1f us, e2e %.3f G/s" % (
            name, d["value"] / 1e9, r["launch_us"], r["frac"], r["replicated_head"]["sync_us_mean"], d["e2e"]["value"] / 1e9))
    except Exception as e:
        print(name, "unreadable:", e)
PY
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_headline.csv python bench.py --only --steps 4 --warmup 3 > $O/b_under_ncu.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_sgd_pipe -s 4 -c 1 -o $O/prof_sgd_pipe python bench.py --only --steps 4 --warmup 3 > $O/ncu1.log 2>&1)
(NRC_SGD_PIPE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_sgd_stream -s 4 -c 1 -o $O/prof_sgd_stream python bench.py --only --steps 4 --warmup 3 > $O/ncu1b.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_csr_fast -s 4 -c 1 -o $O/prof_spmm python bench.py --workload lightgcn-gowalla --only --steps 3 --warmup 3 > $O/ncu2.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_candidate -s 2 -c 1 -o $O/prof_tc python bench.py --workload eval-synth --only --steps 1 --warmup 3 > $O/ncu3.log 2>&1)
(NRC_EVAL_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ncf_scores_tile -c 1 -o $O/prof_ncf_scores python profiles/dbg_round2.py > $O/ncu5.log 2>&1)
cat $O/rc.log; tail -n 6 $O/pytest_pipe.log; tail -n 6 $O/pytest_all.log; tail -3 $O/smoke.log; cat $O/sgd_forms.log; tail -c 400 $O/bench_n1.err; ls -la $O
