#!/bin/bash
# 1-GPU session: the whole GPU suite with the SpectralCF / split / SBPR / APR additions, the reference arm with the
# final workload text, launch list of the final default command.
O=gpurun_out/${1:-r2j}; mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_extras.py -q -m gpu -x 2>&1 | tail -40) > $O/pytest_extras.log 2>&1
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40) > $O/pytest_all.log 2>&1
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; echo "rc smoke $?" >> $O/rc.log
(timeout 600 python bench.py --impl reference > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err); echo "rc bench_ref $?" >> $O/rc.log
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_headline.csv python bench.py --only --steps 4 --warmup 3 > $O/b_under_ncu.log 2>&1)
(timeout 300 python profiles/dbg_spectral.py) > $O/dbg_spectral.log 2>&1
cat $O/rc.log; tail -n 30 $O/pytest_extras.log; tail -n 8 $O/pytest_all.log; tail -3 $O/smoke.log; cat $O/dbg_spectral.log
