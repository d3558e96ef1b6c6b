#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the neurec_b200 hot path.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one training batch (one `sess.run((loss, optimizer), feed_dict)` of the reference's
train_model) through the fused sm_100a kernels.  One JSON line is printed by rank 0:

  value     whole-job triplets/s with the epoch's (user, item, neg|label) arrays already
            resident in HBM when the timed region starts (device sampler ran before it)
  e2e       the same metric through the reference-facing per-step C-ABI call with HOST
            buffers: per step H2D of the batch ids from pinned memory, both kernels, D2H of
            the loss and a stream sync (the analogue of sess.run returning the loss)
  eval      users/s of the full-catalogue evaluator (predict + mask + top-K + 5 metrics)
  roofline  dominant kernel's algorithmic bytes / its CUDA-event launch time vs the measured
            HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's CPU path (oracle/ref_port.py: real compiled reference pieces
            from oracle/_ref where they exist + the numpy restatement of the TF-1.12 step)
            timed on this box's host cores on a bounded sample

Multi-GPU: the training path of these table sizes does not shard (a 2 MB model with ~5 us
steps; see DESIGN.md "replicas only"), so --gpus N runs N independent replicas (weak scaling);
the evaluator shards users across ranks with one all-reduce of the metric sums.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = ("bprmf-ml100k", "neumf-ml100k", "lightgcn-gowalla")
DEFAULT_WORKLOAD = "bprmf-ml100k"


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def load_ml100k():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_split.npz"))
    return {"num_users": int(z["num_users"]), "num_items": int(z["num_items"]),
            "train_indptr": z["train_indptr"].astype(np.int64),
            "train_indices": z["train_indices"].astype(np.int32),
            "test_indptr": z["test_indptr"].astype(np.int64),
            "test_indices": z["test_indices"].astype(np.int32)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index=0, period=0.01):
        self.samples, self.period, self.index = [], period, index
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                self.samples.append((time.perf_counter(), clk, rs, util))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()

    def stop(self):
        if self.ok:
            self._stop.set()
            self.t.join()

    def summary(self, t0=None, t1=None):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": ["nvml unavailable"]}
        nv = self.nv
        sel = [s for s in self.samples if (t0 is None or s[0] >= t0) and (t1 is None or s[0] <= t1)]
        where = "timed region"
        if len(sel) < 3:
            sel, where = self.samples, "whole run (timed region shorter than the sampling period)"
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        bits = 0
        for s in sel:
            bits |= s[2]
        reasons = [k for k, v in names.items() if bits & v]
        busy = [s[1] for s in sel if s[3] > 0] or [s[1] for s in sel]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max, "reasons": reasons,
                "samples": len(sel), "window": where}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


_FLUSH = None


def flush_l2():
    """Write a 256 MiB buffer (> 126 MB L2) so the next kernel starts with a cold L2."""
    import torch
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _FLUSH.fill_(1)


# ----------------------------------------------------------------------------------------
# workload: BPRMF on ml-100k (BASELINE.json configs[0]; conf/MF.properties)
# ----------------------------------------------------------------------------------------
class BprmfMl100k:
    name = "bprmf-ml100k"
    describe = "BPRMF on ml-100k, dim=64, conf/MF.properties (bs 512, adam 1e-3, bpr, reg 0)"
    dim, batch, lr, reg, loss, opt, pairwise, neg_num = 64, 512, 1e-3, 0.0, "bpr", "adam", True, 1
    hyper = [1e-3, 0.9, 0.999, 1e-8]

    def __init__(self, rank=0):
        self.d = load_ml100k()
        self.rank = rank
        d = self.d
        self.users_of_pos = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
        self.n_pos = len(self.users_of_pos)
        self.steps_per_epoch = (self.n_pos + self.batch - 1) // self.batch
        rs = np.random.RandomState(2017 + rank)
        self.U0 = (rs.randn(d["num_users"], self.dim) * 0.01).astype(np.float32)   # normal(0, .01)
        self.V0 = (rs.randn(d["num_items"], self.dim) * 0.01).astype(np.float32)

    # ---------------------------------------------------------------- device state
    def setup_device(self):
        import torch
        from oracle import tf_math  # only for the fp32 lr_t schedule helper (host arithmetic)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        d = self.d
        self.tp, self.ti = dev(d["train_indptr"]), dev(d["train_indices"])
        self.sp, self.si = dev(d["test_indptr"]), dev(d["test_indices"])
        self.dU, self.dV = dev(self.U0), dev(self.V0)
        z = torch.zeros_like
        self.gU, self.gV = z(self.dU), z(self.dV)
        self.mU, self.vU, self.mV, self.vV = z(self.dU), z(self.dU), z(self.dV), z(self.dV)
        self.tU = torch.zeros(d["num_users"], dtype=torch.int32, device="cuda")
        self.tV = torch.zeros(d["num_items"], dtype=torch.int32, device="cuda")
        self.d_users_of_pos = dev(self.users_of_pos)
        self.stamp = 1
        self.lr_t = tf_math.adam_lr_t(self.lr, 1 << 16)
        self.t = 0

    def device_epoch_arrays(self, n_steps, epoch):
        """Device sampler + shuffle for n_steps batches (several epochs if needed)."""
        import torch
        from neurec_b200 import ops
        need = n_steps * self.batch
        us, ps, ns = [], [], []
        e = 0
        while need > 0:
            neg = ops.sample_negatives(self.tp, self.ti, self.d_users_of_pos, 1, self.d["num_items"],
                                       2018, epoch + e)[:, 0]
            perm = torch.randperm(self.n_pos, device="cuda")
            take = min(need, self.n_pos)
            perm = perm[:take]
            us.append(self.d_users_of_pos[perm]); ps.append(self.ti[perm]); ns.append(neg[perm])
            need -= take; e += 1
        cat = lambda xs: torch.cat(xs).contiguous()
        return cat(us), cat(ps), cat(ns)

    def run_steps_device(self, users, pos, neg, n_steps):
        import torch
        from neurec_b200 import ops
        n = min(users.numel(), n_steps * self.batch)
        step_loss = torch.empty(n_steps, device="cuda")
        ops.mf_train_epoch(self.dU, self.dV, users[:n], pos[:n], neg[:n], self.batch, True, self.loss,
                           self.reg, self.opt, self.lr_t[self.t:self.t + n_steps], self.hyper, self.gU,
                           self.gV, self.tU, self.tV, self.mU, self.vU, self.mV, self.vV, self.stamp,
                           step_loss)
        self.stamp += n_steps; self.t += n_steps
        return step_loss, 2 * n_steps

    # ---------------------------------------------------------------- e2e (host buffers)
    def run_steps_e2e(self, h_users, h_pos, h_neg, n_steps):
        """Per step: pinned host ids -> H2D -> grad kernel -> optimizer kernel -> loss D2H."""
        import ctypes
        import torch
        from neurec_b200 import _lib
        lib = _lib.load()
        staging = torch.empty(3 * self.batch + 4, dtype=torch.int32, device="cuda")
        loss_h = torch.zeros(1, dtype=torch.float32).pin_memory()
        hyper = np.array(self.hyper, dtype=np.float32)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        total = 0.0
        bs = self.batch
        for s in range(n_steps):
            hyper[0] = self.lr_t[self.t]
            o = s * bs
            rc = lib.nrc_mf_train_step_host(
                p(self.dU), p(self.dV), self.d["num_users"], self.d["num_items"], self.dim,
                ctypes.c_void_p(h_users.data_ptr() + 4 * o), ctypes.c_void_p(h_pos.data_ptr() + 4 * o),
                ctypes.c_void_p(h_neg.data_ptr() + 4 * o), bs, 1, _lib.LOSS_IDS[self.loss], self.reg,
                _lib.OPT_IDS[self.opt], hyper.ctypes.data, p(self.gU), p(self.gV), p(self.tU), p(self.tV),
                p(self.mU), p(self.vU), p(self.mV), p(self.vV), self.stamp, p(staging), p(loss_h), st)
            _lib.check(rc)
            total += float(loss_h[0])
            self.stamp += 1; self.t += 1
        return total, 3 * 4 * bs, 4

    # ---------------------------------------------------------------- evaluator
    def run_eval(self, users):
        from neurec_b200 import ops
        res = ops.eval_mf(self.dU, self.dV, users, self.tp, self.ti, self.sp, self.si,
                          ["Precision", "Recall", "NDCG", "MAP", "MRR"], 20)
        return ops.mean_rows(res)

    # ---------------------------------------------------------------- roofline inputs
    def algorithmic_bytes(self):
        """per triplet gather 3 rows + ids = 12d+12 B; TF-faithful Adam moves
        (U+I)*d*4 B * 3 arrays * (read+write) per step."""
        rows = self.d["num_users"] + self.d["num_items"]
        grad = self.batch * (12 * self.dim + 12)
        adam = rows * self.dim * 4 * 3 * 2
        return {"mf_pairwise_grad_kernel": grad, "opt_apply_kernel": adam}

    # ---------------------------------------------------------------- CPU reference path
    def cpu_reference(self, n_steps, threads):
        """The reference's CPU path for n_steps batches: real sampler pieces + numpy TF step."""
        from oracle import ref_port, tf_math
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        np.random.seed(2018)
        sampler = ref_port.PairwiseSamplerPort(train_dict, d["num_items"], 1, self.batch, True)
        tr = tf_math.MFTrainer(self.U0, self.V0, self.opt, self.lr, self.loss, self.reg, True)
        done, t0 = 0, time.perf_counter()
        while done < n_steps:
            for bu, bp, bn in sampler:
                tr.step(np.asarray(bu, np.int32), np.asarray(bp, np.int32), np.asarray(bn, np.int32))
                done += 1
                if done >= n_steps:
                    break
        dt = time.perf_counter() - t0
        return dt, ref_port.sampler_kind()

    def cpu_eval(self, threads):
        from oracle import ref_port
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        test_dict = ref_port.user_dict(d["test_indptr"], d["test_indices"])
        t0 = time.perf_counter()
        _, parts, impl = ref_port.evaluate(self.U0, self.V0, train_dict, test_dict, [1, 2, 4, 3, 5], 20,
                                           128, threads)
        return time.perf_counter() - t0, len(test_dict), impl


def make_workload(name, rank):
    if name == "bprmf-ml100k":
        return BprmfMl100k(rank)
    raise SystemExit("workload %s is not available in this build" % name)


# ----------------------------------------------------------------------------------------
# arms
# ----------------------------------------------------------------------------------------
def time_dominant_kernels(w, users, pos, neg, n_steps):
    """Per-kernel CUDA-event timing (each launch bracketed by events on the launching stream)."""
    import torch
    from neurec_b200 import ops
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tg, to = [], []
    loss = torch.zeros(1, device="cuda")
    bs = w.batch
    for s in range(n_steps):
        sl = slice(s * bs, (s + 1) * bs)
        a, b, c = ev(), ev(), ev()
        a.record()
        ops.mf_pairwise_grad(w.dU, w.dV, users[sl], pos[sl], neg[sl], w.loss, w.reg, w.gU, w.gV, w.tU, w.tV,
                             w.stamp, loss)
        b.record()
        hyper = np.array(w.hyper, dtype=np.float32); hyper[0] = w.lr_t[w.t]
        # both tables in one launch, exactly as the epoch driver does
        ops.opt_apply_multi(w.opt, [(w.dU, w.gU, w.mU, w.vU, w.tU, False),
                                    (w.dV, w.gV, w.mV, w.vV, w.tV, False)], w.stamp, hyper)
        c.record()
        w.stamp += 1; w.t += 1
        tg.append((a, b)); to.append((b, c))
    torch.cuda.synchronize()
    g = float(np.mean([x.elapsed_time(y) for x, y in tg])) * 1e-3
    o = float(np.mean([x.elapsed_time(y) for x, y in to])) * 1e-3
    return {"mf_pairwise_grad_kernel": g, "opt_apply_kernel": o}


def run_ours(args):
    import torch
    rank, world, local = dist_setup(args.gpus)
    w = make_workload(args.workload, rank)
    w.setup_device()
    K, W = args.steps, max(args.warmup, 3)
    clocks = ClockSampler(local)
    clocks.start()

    users, pos, neg = w.device_epoch_arrays(K + W, epoch=0)
    # ---- value: device-resident steps -------------------------------------------------
    w.run_steps_device(users, pos, neg, W)
    barrier(world)
    flush_l2()
    barrier(world)
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o = W * w.batch
    e0.record()
    _, launches = w.run_steps_device(users[o:], pos[o:], neg[o:], K)
    e1.record()
    barrier(world)
    t_wall1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    value = world * K * w.batch / (ms * 1e-3)

    # ---- e2e: per-step host buffers ---------------------------------------------------
    h_users, h_pos, h_neg = (t.cpu().pin_memory() for t in (users, pos, neg))
    w.run_steps_e2e(h_users, h_pos, h_neg, W)
    barrier(world)
    flush_l2()
    barrier(world)
    t0 = time.perf_counter()
    _, h2d, d2h = w.run_steps_e2e(h_users[o:], h_pos[o:], h_neg[o:], K)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    barrier(world)
    e2e_value = world * K * w.batch / e2e_s

    # ---- evaluator: users sharded over ranks -------------------------------------------
    nu = w.d["num_users"]
    mine = torch.arange(rank, nu, world, dtype=torch.int32, device="cuda")
    w.run_eval(mine)
    barrier(world)
    flush_l2()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        w.run_eval(mine)
    b.record()
    barrier(world)
    eval_ms = max_over_ranks(a.elapsed_time(b), world) / reps
    clocks.stop()

    out = None
    if rank == 0:
        kt = time_dominant_kernels(w, users, pos, neg, min(K, 64))
        ab = w.algorithmic_bytes()
        dom = max(kt, key=lambda k: kt[k])
        peak, peak_src = measured_peaks()
        achieved = ab[dom] / kt[dom] / 1e9
        # CPU baseline on a bounded sample
        threads = os.cpu_count() or 1
        n_cpu = min(w.steps_per_epoch, 157)
        dt, skind = w.cpu_reference(n_cpu, threads)
        cpu_value = n_cpu * w.batch / dt
        dte, n_eval_users, eimpl = w.cpu_eval(threads)
        out = {
            "metric": "triplets/sec", "value": value, "unit": "triplets/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "ml-100k ratio-0.8 split of the reference (tests/golden/ml100k_split.npz), "
                    "normal(0,0.01) random-init tables, device Philox negatives",
            "config": {"workload": w.describe, "global_batch": w.batch * world, "dim": w.dim,
                       "optimizer": "adam (TF-1.12 dense-over-table semantics)",
                       "parallelism": "replicas x%d (training does not shard at this size)" % world,
                       "l2": "flushed (256 MiB write) before the timed region; the K dependent "
                             "steps then run back-to-back as in training (working set 2.7 MB)"},
            "e2e": {"value": e2e_value, "unit": "triplets/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3 / K},
            "gpu_launches": launches,
            "eval": {"metric": "eval users/sec", "value": nu / (eval_ms * 1e-3), "unit": "users/s",
                     "users": nu, "items": w.d["num_items"], "top_k": 20, "metrics": 5,
                     "ms": eval_ms, "sharding": "users over %d rank(s)" % world,
                     "cpu": {"value": n_eval_users / dte, "unit": "users/s", "kind":
                             "reference" if eimpl == "reference" else "port",
                             "what": "np.matmul predict + python mask loop + evaluate.h top-K/metrics, "
                                     "test_batch_size 128, %d threads" % threads}},
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "bytes_per_launch": ab[dom],
                         "launch_us": kt[dom] * 1e6,
                         "all_kernels_us": {k: v * 1e6 for k, v in kt.items()},
                         "note": "tables (2.7 MB with Adam state) are L2-resident: the honest bound "
                                 "here is launch latency, not HBM"},
            "cpu_baseline": {"value": cpu_value, "unit": "triplets/s", "cores": threads,
                             "kind": "port", "sample": "%d steps of %d (one ml-100k epoch): reference "
                             "sampler (%s random_choice) + numpy restatement of the TF-1.12 BPR/Adam "
                             "step" % (n_cpu, w.batch, skind)},
            "clocks": clocks.summary(t_wall0, t_wall1),
        }
    barrier(world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def run_reference(args):
    """The reference's own CPU path on this box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload, 0)
    threads = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    # bounded sample made of WHOLE epochs (the reference draws an epoch's negatives up front,
    # so a partial epoch would overcharge it): 1..3 epochs of steps_per_epoch batches
    spe = w.steps_per_epoch
    n = spe * min(max(K // spe, 1), 3)
    w.cpu_reference(min(W, 20), threads)
    dt, skind = w.cpu_reference(n, threads)
    value = n * w.batch / dt
    out = {"impl": "reference", "metric": "triplets/sec", "value": value, "unit": "triplets/s",
           "n_gpus": args.gpus, "steps": n, "warmup": W, "ms_per_step": dt * 1e3 / n,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "ml-100k ratio-0.8 split of the reference, same init tables",
           "config": {"workload": w.describe, "global_batch": w.batch},
           "cpu_baseline": {"value": value, "unit": "triplets/s", "cores": threads, "kind": "port",
                            "sample": "%d steps: reference sampler (%s random_choice) + numpy "
                                      "restatement of the TF-1.12 step (TensorFlow 1.12 is not "
                                      "installable offline)" % (n, skind)},
           "e2e": {"value": value, "unit": "triplets/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=157)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=WORKLOADS)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
