#!/bin/bash
# 1-GPU session: the rank-3 plug-ins on their real datasets, ncu captures of the new kernels.
O=gpurun_out/${1:-r2r}; mkdir -p $O
(timeout 500 python profiles/dbg_extras_real.py) > $O/extras_real.log 2>&1; echo "rc real $?" >> $O/rc.log
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 8 -c 1 -o $O/prof_dense_gemm python profiles/dbg_spectral.py > $O/ncu_a.log 2>&1)
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:mf_bpr_lazy_adam -s 3 -c 1 -o $O/prof_lazy_adam python bench.py --only --steps 4 --warmup 3 > $O/ncu_b.log 2>&1)
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:sbpr_grad_kernel -s 2 -c 1 -o $O/prof_sbpr_grad python -m pytest tests/test_gpu_extras.py -q -m gpu -k "sbpr_train_epoch_vs_oracle and adam" > $O/ncu_c.log 2>&1)
cat $O/rc.log $O/extras_real.log; ls $O
