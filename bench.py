#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the neurec_b200 hot path.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference] [--only]

A "step" is one training batch (one `sess.run((loss, optimizer), feed_dict)` of the reference's
train_model) through the fused sm_100a kernels.  Rank 0 prints ONE JSON line:

  value     whole-job samples/s ("triplets/s"; a pointwise sample counts as one triplet,
            SURVEY.md 8d) with the epoch's id arrays already resident in HBM when the timed
            region starts (the device sampler ran before it)
  e2e       the same metric through the reference-facing per-step calls with HOST buffers: per
            step H2D of the batch ids/labels from pinned memory, the step's kernels, D2H of the
            loss and a stream sync (the analogue of sess.run returning the loss)
  eval      users/s of the full-catalogue evaluator (predict + mask + top-K + 5 metrics)
  roofline  dominant kernel: algorithmic bytes per launch / its CUDA-event launch time (graph
            replay of that kernel alone) vs the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the reference's CPU path (oracle/ref_port.py: the real compiled reference
            pieces from oracle/_ref where they exist + the numpy restatement of the TF-1.12
            step) timed on this box's host cores on a bounded sample
  others    (N=1, unless --only) the same measurements for the other single-GPU configs

Workloads (BASELINE.json configs): neumf-ml100k (configs[1], default), bprmf-ml100k
(configs[0]), lightgcn-gowalla (configs[2], synthetic gowalla-shaped graph).

Multi-GPU: training of these table sizes does not shard (2-20 MB models with microsecond
steps; DESIGN.md "replicas only"), so --gpus N runs N independent replicas (weak scaling); the
evaluator shards users across ranks and all-gathers the per-user rows.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = ("neumf-ml100k", "bprmf-ml100k", "lightgcn-gowalla")
DEFAULT_WORKLOAD = "neumf-ml100k"
METRICS = ["Precision", "Recall", "NDCG", "MAP", "MRR"]


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def load_ml100k():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_split.npz"))
    return {"name": "ml-100k", "num_users": int(z["num_users"]), "num_items": int(z["num_items"]),
            "train_indptr": z["train_indptr"].astype(np.int64),
            "train_indices": z["train_indices"].astype(np.int32),
            "test_indptr": z["test_indptr"].astype(np.int64),
            "test_indices": z["test_indices"].astype(np.int32)}


def synth_gowalla(seed=7):
    """Synthetic graph with gowalla's shape (U=29 858, I=40 981, ~810 k train and
    ~217 k test interactions, power-law item popularity, user degrees 8..~800)."""
    rs = np.random.RandomState(seed)
    nu, ni = 29858, 40981
    deg = np.clip((8 + rs.pareto(1.35, nu) * 9).astype(np.int64), 8, 811)
    deg = (deg * (810128 / deg.sum())).astype(np.int64).clip(6, 811)
    pop = 1.0 / np.power(np.arange(1, ni + 1) + 50.0, 0.75)   # head capped like gowalla (max item degree ~1.4 k)
    pop = pop[rs.permutation(ni)]
    pop /= pop.sum()
    tot = int((deg * 1.45).sum())
    draws = rs.choice(ni, size=tot, p=pop).astype(np.int32)
    off = np.concatenate([[0], np.cumsum((deg * 1.45).astype(np.int64))])
    tr_rows, te_rows = [], []
    for u in range(nu):
        it = np.unique(draws[off[u]:off[u + 1]])
        rs.shuffle(it)
        k = max(1, int(round(len(it) * 0.79)))
        tr_rows.append(np.sort(it[:k])); te_rows.append(np.sort(it[k:]) if len(it) > k else np.sort(it[:1]))

    def csr(rows):
        ptr = np.zeros(nu + 1, np.int64)
        ptr[1:] = np.cumsum([len(r) for r in rows])
        return ptr, np.concatenate(rows).astype(np.int32)
    tp, ti = csr(tr_rows)
    sp_, si = csr(te_rows)
    return {"name": "gowalla-shaped synthetic", "num_users": nu, "num_items": ni, "train_indptr": tp,
            "train_indices": ti, "test_indptr": sp_, "test_indices": si}


def profiled_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.isfile(p):
        return json.load(open(p)).get(kernel)
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index=0, period=0.005):
        self.samples, self.period = [], period
        self._stop = threading.Event()
        self.ok, self.max = False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(self.h),
                                     nv.nvmlDeviceGetUtilizationRates(self.h).gpu))
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

    def summary(self, windows):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": ["nvml unavailable"]}
        nv = self.nv
        sel = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        where = "timed regions"
        if len(sel) < 3:
            sel, where = self.samples, "whole run (timed regions shorter than the sampling period)"
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        bits = 0
        for s in sel:
            bits |= s[2]
        busy = [s[1] for s in sel if s[3] > 0] or [s[1] for s in sel]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max,
                "reasons": [k for k, v in names.items() if bits & v], "samples": len(sel), "window": where}


def dist_setup():
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


def graph_time(fn, reps=40, rounds=5):
    """Average device time of one `fn()` launch group: `reps` calls captured in a CUDA graph and
    replayed (no host launch overhead), CUDA events on the replay stream, best of `rounds`."""
    import torch
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # thread-local capture mode: the NCCL watchdog thread of a multi-rank run must stay free to
        # query its events while this thread captures
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(rounds):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            g.replay()
            b.record(s)
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
    torch.cuda.current_stream().wait_stream(s)
    return best * 1e-3 / reps


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ----------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------
class Workload:
    """Common plumbing: data, device sampler/shuffle, e2e staging, evaluator, CPU reference."""
    pairwise = True
    neg_num = 1
    eval_k = 20

    def __init__(self, data, rank):
        self.d, self.rank = data, rank
        d = data
        self.users_of_pos = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
        self.n_pos = len(self.users_of_pos)
        self.n_samples = self.n_pos * (1 if self.pairwise else self.neg_num + 1)
        self.steps_per_epoch = (self.n_samples + self.batch - 1) // self.batch

    def setup_common(self):
        import torch
        d = self.d
        self.tp, self.ti = dev(d["train_indptr"]), dev(d["train_indices"])
        self.sp, self.si = dev(d["test_indptr"]), dev(d["test_indices"])
        self.d_users_of_pos = dev(self.users_of_pos)
        self.staging = torch.empty(3 * self.batch + 4, dtype=torch.int32, device="cuda")
        self.loss_pin = torch.zeros(4, dtype=torch.float32).pin_memory()
        self.lr_sched = self._lr_schedule(1 << 17)
        self.t = 0
        self.stamp = 1

    def _lr_schedule(self, n):
        out = np.empty(n, np.float32)
        p1, p2, one, lr = np.float32(0.9), np.float32(0.999), np.float32(1), np.float32(self.lr)
        for s in range(n):
            out[s] = lr * np.sqrt(one - p2) / (one - p1)
            p1 = np.float32(p1 * np.float32(0.9)); p2 = np.float32(p2 * np.float32(0.999))
        return out

    def epoch_arrays(self, n_steps, epoch=0):
        """Device sampler + shuffle for n_steps batches (wraps over epochs if needed)."""
        import torch
        from neurec_b200 import ops
        need = n_steps * self.batch
        us, its, th = [], [], []
        e = 0
        while need > 0:
            neg = ops.sample_negatives(self.tp, self.ti, self.d_users_of_pos, self.neg_num,
                                       self.d["num_items"], 2018, epoch + e)
            if self.pairwise:
                u, i, t = self.d_users_of_pos, self.ti, neg[:, 0]
            else:
                u = self.d_users_of_pos.repeat(self.neg_num + 1)
                i = torch.cat([self.ti, neg.t().reshape(-1)])
                t = torch.cat([torch.ones(self.n_pos, device="cuda"),
                               torch.zeros(self.n_pos * self.neg_num, device="cuda")])
            perm = torch.randperm(u.numel(), device="cuda")[:min(need, u.numel())]
            us.append(u[perm]); its.append(i[perm]); th.append(t[perm])
            need -= perm.numel(); e += 1
        cat = lambda xs: torch.cat(xs).contiguous()
        return cat(us), cat(its), cat(th)

    def stage(self, h_arrays, off):
        """H2D of one batch of (users, items, third) from pinned host memory."""
        from neurec_b200 import _lib
        import torch
        lib = _lib.load()
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = [ctypes.c_void_p(h.data_ptr() + 4 * off) for h in h_arrays]
        _lib.check(lib.nrc_stage_batch_host(p[0], p[1], p[2], self.batch, ctypes.c_void_p(self.staging.data_ptr()), st))
        b = self.batch
        import torch as T
        third = self.staging[2 * b:3 * b]
        if not self.pairwise:
            third = third.view(T.float32)
        return self.staging[:b], self.staging[b:2 * b], third

    def fetch(self, dev_tensor, count):
        from neurec_b200 import _lib
        import torch
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().nrc_fetch_host(ctypes.c_void_p(dev_tensor.data_ptr()),
                                              ctypes.c_void_p(self.loss_pin.data_ptr()), count, st))
        return float(self.loss_pin[0])

    RING = 16

    def _capture_step(self, lib, r, copy_in, main, copy_out):
        """Issue step r of the ring into the current capture: H2D on `copy_in`, kernels on `main`,
        loss D2H on `copy_out` (the three may be the same stream)."""
        import torch
        from neurec_b200 import _lib
        vp = ctypes.c_void_p
        b = self.batch
        stag = self.stag_ring[r]
        _lib.check(lib.nrc_graph_stage_async(vp(self.pins[r].data_ptr()), vp(stag.data_ptr()), (3 * b + 1) * 4,
                                             vp(copy_in.cuda_stream)))
        if copy_in is not main:
            _lib.check(lib.nrc_graph_depend(vp(copy_in.cuda_stream), vp(main.cuda_stream)))
        _lib.check(lib.nrc_opt_set_lr_source(vp(stag.data_ptr() + 12 * b)))
        third = stag[2 * b:3 * b]
        if not self.pairwise:
            third = third.view(torch.float32)
        t_keep, loss_keep = self.t, self.step_loss
        self.step_loss = self.loss_ring[r]
        with torch.cuda.stream(main):
            self.step_on_staged(stag[:b], stag[b:2 * b], third)
        self.t, self.step_loss = t_keep, loss_keep
        _lib.check(lib.nrc_opt_set_lr_source(None))
        if copy_out is not main:
            _lib.check(lib.nrc_graph_depend(vp(main.cuda_stream), vp(copy_out.cuda_stream)))
        _lib.check(lib.nrc_graph_fetch_async(vp(self.loss_ring[r].data_ptr()), vp(self.loss_pins[r].data_ptr()),
                                             self.loss_count, vp(copy_out.cuda_stream)))

    def build_step_graph(self):
        """Captured training steps for the host-facing path.  Per ring slot r: a pinned block, a device
        staging block, a device + pinned loss slot and a single-step graph (H2D, the step's kernels,
        loss D2H).  Plus ONE burst graph of RING consecutive steps whose H2D chain and loss-D2H chain
        are captured on two side streams: inside a burst the copy of step s+1 overlaps the kernels of
        step s (nrc_graph_run_steps)."""
        import torch
        from neurec_b200 import _lib
        lib = _lib.load()
        b, R = self.batch, self.RING
        self.e2e_stream = torch.cuda.Stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self.pins = [torch.zeros(3 * b + 4, dtype=torch.int32).pin_memory() for _ in range(R)]
        self.loss_pins = [torch.zeros(max(self.loss_count, 1)).pin_memory() for _ in range(R)]
        self.stag_ring = torch.zeros((R, 3 * b + 4), dtype=torch.int32, device="cuda")
        self.loss_ring = torch.zeros((R, 16), device="cuda")
        self.graphs = []
        torch.cuda.synchronize()
        vp = ctypes.c_void_p
        main = self.e2e_stream
        st = vp(main.cuda_stream)
        for r in range(R):
            g = ctypes.c_void_p()
            _lib.check(lib.nrc_graph_capture_begin(st))
            self._capture_step(lib, r, main, main, main)
            _lib.check(lib.nrc_graph_capture_end(st, ctypes.byref(g)))
            self.graphs.append(g)
        self.burst = ctypes.c_void_p()
        _lib.check(lib.nrc_graph_capture_begin(st))
        _lib.check(lib.nrc_graph_depend(st, vp(s_in.cuda_stream)))        # fork: both side streams join the capture
        _lib.check(lib.nrc_graph_depend(st, vp(s_out.cuda_stream)))
        for r in range(R):
            self._capture_step(lib, r, s_in, main, s_out)
        _lib.check(lib.nrc_graph_depend(vp(s_in.cuda_stream), st))        # join
        _lib.check(lib.nrc_graph_depend(vp(s_out.cuda_stream), st))
        _lib.check(lib.nrc_graph_capture_end(st, ctypes.byref(self.burst)))
        self.graph, self.pin, self.loss_pin = self.graphs[0], self.pins[0], self.loss_pins[0]
        self.c_graphs = (ctypes.c_void_p * R)(*[g.value for g in self.graphs])
        self.c_pins = (ctypes.c_void_p * R)(*[p.data_ptr() for p in self.pins])
        self.c_loss = (ctypes.c_void_p * R)(*[p.data_ptr() for p in self.loss_pins])
        torch.cuda.synchronize()

    def run_steps_e2e(self, h_arrays, n_steps):
        """Steps from HOST arrays: the host stages RING batches into the pinned ring and launches the
        burst graph (per step: H2D node, the step's kernels, loss D2H node; copies of step s+1 overlap
        the kernels of step s), waits, reads the RING losses; leftover steps use the single-step graphs."""
        from neurec_b200 import _lib
        lib = _lib.load()
        if not hasattr(self, "graphs"):
            self.build_step_graph()
        st = ctypes.c_void_p(self.e2e_stream.cuda_stream)
        lr = np.ascontiguousarray(self.lr_sched[self.t:self.t + n_steps], dtype=np.float32)
        total = ctypes.c_double(0.0)
        vp = ctypes.c_void_p
        _lib.check(lib.nrc_graph_run_steps(self.burst, self.c_graphs, self.RING, vp(h_arrays[0].data_ptr()),
                                           vp(h_arrays[1].data_ptr()), vp(h_arrays[2].data_ptr()), self.batch,
                                           lr.ctypes.data_as(ctypes.c_void_p), n_steps, self.c_pins, self.c_loss,
                                           self.loss_count, ctypes.byref(total), st))
        self.t += n_steps
        return total.value, 3 * 4 * self.batch + 4, 4 * self.loss_count

    def run_steps_e2e_sync(self, h_arrays, n_steps):
        """The strictly synchronous variant (what `sess.run` per batch does): stage, launch, wait for
        the loss, every step."""
        from neurec_b200 import _lib
        lib = _lib.load()
        if not hasattr(self, "graphs"):
            self.build_step_graph()
        st = ctypes.c_void_p(self.e2e_stream.cuda_stream)
        pin = ctypes.c_void_p(self.pin.data_ptr())
        ptrs = [h.data_ptr() for h in h_arrays]
        loss_np = self.loss_pin.numpy()
        total, b = 0.0, self.batch
        for s in range(n_steps):
            o = 4 * s * b
            rc = lib.nrc_graph_step(self.graph, ctypes.c_void_p(ptrs[0] + o), ctypes.c_void_p(ptrs[1] + o),
                                    ctypes.c_void_p(ptrs[2] + o), b, float(self.lr_sched[self.t]), pin, st)
            if rc:
                _lib.check(rc)
            self.t += 1
            total += float(loss_np[0])
        return total, 3 * 4 * self.batch + 4, 4 * self.loss_count

    def cpu_eval(self, threads, U, V, max_users=None):
        from oracle import ref_port
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        test_dict = ref_port.user_dict(d["test_indptr"], d["test_indices"])
        if max_users is not None and len(test_dict) > max_users:
            keys = list(test_dict.keys())[:max_users]
            test_dict = {k: test_dict[k] for k in keys}
        best = None
        for th in sorted({8, min(threads, 32)}):     # reference default num_thread=8 vs more threads
            t0 = time.perf_counter()
            _, parts, impl = ref_port.evaluate(U, V, train_dict, test_dict, [1, 2, 4, 3, 5], self.eval_k, 128, th)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, th, parts, impl)
        return best[0], len(test_dict), best[3], best[1]


class BprmfMl100k(Workload):
    name = "bprmf-ml100k"
    describe = "BPRMF on ml-100k, dim=64, conf/MF.properties (bs 512, adam 1e-3, bpr, reg 0)"
    dim, batch, lr, reg, loss, opt = 64, 512, 1e-3, 0.0, "bpr", "adam"
    hyper = [1e-3, 0.9, 0.999, 1e-8]
    loss_count = 1
    launches_per_step = 2

    def __init__(self, rank=0):
        super().__init__(load_ml100k(), rank)
        rs = np.random.RandomState(2017 + rank)
        self.U0 = (rs.randn(self.d["num_users"], self.dim) * 0.01).astype(np.float32)
        self.V0 = (rs.randn(self.d["num_items"], self.dim) * 0.01).astype(np.float32)

    def setup_device(self):
        import torch
        self.setup_common()
        self.dU, self.dV = dev(self.U0), dev(self.V0)
        z = torch.zeros_like
        self.gU, self.gV = z(self.dU), z(self.dV)
        self.mU, self.vU, self.mV, self.vV = z(self.dU), z(self.dU), z(self.dV), z(self.dV)
        self.tU = torch.zeros(self.d["num_users"], dtype=torch.int32, device="cuda")
        self.tV = torch.zeros(self.d["num_items"], dtype=torch.int32, device="cuda")
        self.step_loss = torch.zeros(1 << 16, device="cuda")

    def run_steps_device(self, arrays, n_steps):
        from neurec_b200 import ops
        u, i, t = arrays
        n = min(u.numel(), n_steps * self.batch)
        ops.mf_train_epoch(self.dU, self.dV, u[:n], i[:n], t[:n], self.batch, True, self.loss, self.reg,
                           self.opt, self.lr_sched[self.t:self.t + n_steps], self.hyper, self.gU, self.gV,
                           self.tU, self.tV, self.mU, self.vU, self.mV, self.vV, self.stamp, self.step_loss)
        self.stamp += n_steps; self.t += n_steps
        return self.launches_per_step * n_steps

    def step_on_staged(self, u, i, t):
        self.run_steps_device((u, i, t), 1)
        return self.step_loss

    def eval_tables(self):
        return self.dU, self.dV

    def kernels(self, arrays):
        from neurec_b200 import ops
        import torch
        u, i, t = (a[:self.batch] for a in arrays)
        loss = torch.zeros(1, device="cuda")
        hyper = list(self.hyper)
        grad = lambda: ops.mf_pairwise_grad(self.dU, self.dV, u, i, t, self.loss, self.reg, self.gU, self.gV,
                                            self.tU, self.tV, 7, loss)
        opt = lambda: ops.opt_apply_multi(self.opt, [(self.dU, self.gU, self.mU, self.vU, self.tU, False),
                                                     (self.dV, self.gV, self.mV, self.vV, self.tV, False)], 7, hyper)
        rows = self.d["num_users"] + self.d["num_items"]
        return {"mf_pairwise_grad_kernel": (grad, self.batch * (12 * self.dim + 12),
                                            "512 triplets x (3 rows of 64 f32 + 3 ids)"),
                "opt_apply_kernel": (opt, rows * self.dim * 4 * 3 * 2,
                                     "TF-faithful Adam: (U+I)*d*4 B x {var,m,v} x read+write")}

    def cpu_reference(self, n_steps):
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
        return time.perf_counter() - t0, ref_port.sampler_kind()

    def cpu_tables(self):
        return self.U0, self.V0


class NeumfMl100k(Workload):
    name = "neumf-ml100k"
    describe = ("NeuMF (GMF+MLP) on ml-100k, embedding_size=32, layers [64,32,16], conf/NeuMF.properties "
                "(pointwise cross_entropy, num_neg 4, bs 256, adam 1e-3)")
    pairwise, neg_num = False, 4
    mf_dim, layers, batch, lr, loss, opt = 32, [64, 32, 16], 256, 1e-3, "cross_entropy", "adam"
    hyper = [1e-3, 0.9, 0.999, 1e-8]
    loss_count = 1
    launches_per_step = 3
    KEYS = ("mf_user", "mf_item", "mlp_user", "mlp_item", "dense")

    def __init__(self, rank=0):
        super().__init__(load_ml100k(), rank)
        rs = np.random.RandomState(2017 + rank)
        nu, ni = self.d["num_users"], self.d["num_items"]
        n = lambda r, c: (rs.randn(r, c) * 0.01).astype(np.float32)
        self.P0 = {"mf_user": n(nu, 32), "mf_item": n(ni, 32), "mlp_user": n(nu, 32), "mlp_item": n(ni, 32)}
        dense, inn = [], 64
        for out in self.layers:                      # glorot-uniform kernels, zero biases
            lim = np.sqrt(6.0 / (inn + out))
            dense += [rs.uniform(-lim, lim, inn * out).astype(np.float32), np.zeros(out, np.float32)]
            inn = out
        self.P0["dense"] = np.concatenate(dense)

    def setup_device(self):
        import torch
        from neurec_b200 import ops
        self.setup_common()
        nu, ni = self.d["num_users"], self.d["num_items"]
        self.shape = ops.NcfShape.make(nu, ni, self.mf_dim, self.layers, 1)
        self.P = {k: dev(v) for k, v in self.P0.items()}
        z = lambda D: {k: torch.zeros_like(v) for k, v in D.items()}
        self.G, self.S0, self.S1 = z(self.P), z(self.P), z(self.P)
        self.tU = torch.zeros(nu, dtype=torch.int32, device="cuda")
        self.tI = torch.zeros(ni, dtype=torch.int32, device="cuda")
        self.step_loss = torch.zeros(1 << 16, device="cuda")

    def run_steps_device(self, arrays, n_steps):
        from neurec_b200 import ops
        u, i, t = arrays
        n = min(u.numel(), n_steps * self.batch)
        ops.ncf_train_epoch(self.shape, self.P, u[:n], i[:n], t[:n], self.batch, False, self.loss, 0.0, 0.0,
                            self.opt, self.lr_sched[self.t:self.t + n_steps], self.hyper, self.G, self.S0,
                            self.S1, self.tU, self.tI, self.stamp, self.step_loss)
        self.stamp += n_steps; self.t += n_steps
        return self.launches_per_step * n_steps

    def step_on_staged(self, u, i, t):
        self.run_steps_device((u, i, t), 1)
        return self.step_loss

    def eval_tables(self):
        return None

    def run_eval(self, users):
        """users must be a contiguous id range (what bench.py's sharding hands out)."""
        from neurec_b200 import ops
        a, b = int(users[0].item()), int(users[-1].item()) + 1
        scores = ops.ncf_scores(self.shape, self.P, users)          # NeuMF.predict over all items
        ops.mask_rows(scores, users, self.tp, self.ti)
        ptr = (self.sp[a:b + 1] - self.sp[a]).contiguous()
        idx = self.si[int(self.sp[a].item()):int(self.sp[b].item())].contiguous()
        return ops.eval_score_matrix(scores, ptr, idx, METRICS, self.eval_k)

    def kernels(self, arrays):
        from neurec_b200 import ops
        import torch
        u, i, t = (a[:self.batch] for a in arrays)
        loss = torch.zeros(1, device="cuda")
        grad = lambda: ops.ncf_grad(self.shape, self.P, u, i, t, False, self.loss, 0.0, 0.0, self.G, self.tU,
                                    self.tI, 7, loss)
        variables = [(self.P[k], self.G[k], self.S0[k], self.S1[k], (self.tU if "user" in k else self.tI), False)
                     for k in self.KEYS[:4]] + [(self.P["dense"], self.G["dense"], self.S0["dense"],
                                                 self.S1["dense"], None, True)]
        opt = lambda: ops.opt_apply_multi(self.opt, variables, 7, list(self.hyper))
        n_par = sum(v.numel() for v in self.P.values())
        return {"ncf_sample_kernel+ncf_wgrad_kernel": (grad, self.batch * (4 * 32 * 4 * 2 + 12) + self.P["dense"].numel() * 4 * 2,
                                    "256 samples x (4 rows of 32 f32 gathered + their gradients + ids) + "
                                    "dense weights read + their gradient written"),
                "opt_apply_kernel": (opt, n_par * 4 * 3 * 2,
                                     "TF-faithful Adam over 4 tables + dense: params*4 B x {var,m,v} x R+W")}

    def cpu_reference(self, n_steps):
        from oracle import ref_port, tf_math
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        np.random.seed(2018)
        sampler = ref_port.PointwiseSamplerPort(train_dict, d["num_items"], self.neg_num, self.batch, True)
        tr = tf_math.NCFTrainer(self.P0, 32, self.layers, 1, self.opt, self.lr, self.loss, 0.0, 0.0, False)
        done, t0 = 0, time.perf_counter()
        while done < n_steps:
            for bu, bi, bl in sampler:
                tr.step(np.asarray(bu, np.int32), np.asarray(bi, np.int32), np.asarray(bl, np.float32))
                done += 1
                if done >= n_steps:
                    break
        return time.perf_counter() - t0, ref_port.sampler_kind()

    def cpu_eval(self, threads, U=None, V=None, max_users=None):
        """NeuMF.predict on the CPU = one forward over all items per user (NeuMF.py:163-168)."""
        from oracle import ref_port, tf_math
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        test_dict = ref_port.user_dict(d["test_indptr"], d["test_indices"])
        keys = list(test_dict.keys())[:256]
        test_dict = {k: test_dict[k] for k in keys}
        items = np.arange(d["num_items"])
        pred = lambda bu: np.stack([tf_math.ncf_predict(self.P0, np.full(len(items), u), items, 32, self.layers)
                                    for u in bu])
        t0 = time.perf_counter()
        _, parts, impl = ref_port.evaluate(None, None, train_dict, test_dict, [1, 2, 4, 3, 5], self.eval_k, 128, 8,
                                           predict=pred)
        return time.perf_counter() - t0, len(test_dict), impl, 8


class LightgcnGowalla(Workload):
    name = "lightgcn-gowalla"
    describe = ("LightGCN on a gowalla-shaped graph (29 858 users, 40 981 items), 3 layers dim=64, "
                "conf/LightGCN.properties (bs 1024, adam 0.01, reg 1e-3, adj_type pre)")
    dim, n_layers, batch, lr, reg = 64, 3, 1024, 0.01, 1e-3
    hyper = [0.01, 0.9, 0.999, 1e-8]
    loss_count = 2

    def __init__(self, rank=0):
        super().__init__(synth_gowalla(), rank)
        self.launches_per_step = 2 * self.n_layers + 3
        from neurec_b200.model.general_recommender.LightGCN import bipartite_adjacency   # the product's builder
        d = self.d
        u = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
        A = bipartite_adjacency(u, d["train_indices"], d["num_users"], d["num_items"], "pre", verbose=False)
        self.A = A.tocoo().astype(np.float32).tocsr()          # LightGCN.py:151-154
        self.A.sort_indices()
        rs = np.random.RandomState(2017 + rank)
        n = d["num_users"] + d["num_items"]
        lim = np.sqrt(6.0 / (d["num_users"] + self.dim))
        self.E0 = rs.uniform(-lim, lim, (n, self.dim)).astype(np.float32)

    def setup_device(self):
        import torch
        self.setup_common()
        A = self.A
        self.csr = (dev(A.indptr.astype(np.int64)), dev(A.indices.astype(np.int32)), dev(A.data.astype(np.float32)))
        self.order = dev(np.argsort(-np.diff(A.indptr), kind="stable").astype(np.int32))
        self.e0 = dev(self.E0)
        z = lambda: torch.zeros_like(self.e0)
        self.m, self.v, self.ef, self.gf, self.ge = z(), z(), z(), z(), z()
        self.work = (z(), z())
        self.step_loss = torch.zeros((1 << 14, 2), device="cuda")

    def run_steps_device(self, arrays, n_steps):
        from neurec_b200 import ops
        u, i, t = arrays
        n = min(u.numel(), n_steps * self.batch)
        d = self.d
        ops.lightgcn_train_epoch(self.csr, None, self.order, d["num_users"], d["num_items"], self.n_layers,
                                 self.e0, self.m, self.v, u[:n], i[:n], t[:n], self.batch, self.reg,
                                 self.lr_sched[self.t:self.t + n_steps], self.hyper, self.ef, self.gf, self.ge,
                                 self.work, self.step_loss)
        self.t += n_steps
        return self.launches_per_step * n_steps

    def step_on_staged(self, u, i, t):
        self.run_steps_device((u, i, t), 1)
        return self.step_loss

    def eval_tables(self):
        from neurec_b200 import ops
        ops.lightgcn_propagate(self.csr[0], self.csr[1], self.csr[2], self.order, self.e0, self.n_layers,
                               self.ef, self.work)
        nu = self.d["num_users"]
        return self.ef[:nu].contiguous(), self.ef[nu:].contiguous()

    def kernels(self, arrays):
        from neurec_b200 import ops
        n, nnz = self.A.shape[0], self.A.nnz
        spmm = lambda: ops.spmm_csr(self.csr[0], self.csr[1], self.csr[2], self.e0, row_order=self.order,
                                    y=self.work[0])
        opt = lambda: ops.opt_apply_multi("adam", [(self.e0, self.ge, self.m, self.v, None, True)], 0,
                                          list(self.hyper))
        return {"spmm_csr_kernel": (spmm, nnz * 8 + (n + 1) * 8 + 2 * n * self.dim * 4,
                                    "SURVEY.md 8(d): nnz*(4 B col + 4 B val) + indptr + N*d*4 B read + written"),
                "opt_apply_kernel": (opt, n * self.dim * 4 * 4 * 2, "dense Adam: N*d*4 B x {var,m,v,grad} x R+W")}

    def cpu_reference(self, n_steps):
        from oracle import ref_port, tf_math
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        np.random.seed(2018)
        t0 = time.perf_counter()
        sampler = ref_port.PairwiseSamplerPort(train_dict, d["num_items"], 1, self.batch, True)
        it = iter(sampler)
        first = next(it)                      # pays the epoch's negative sampling
        t_sample = time.perf_counter() - t0
        tr = tf_math.LightGCNTrainer(self.A, self.E0, d["num_users"], self.n_layers, self.lr, self.reg)
        t1 = time.perf_counter()
        done, batch = 0, first
        while done < n_steps:
            tr.step(np.asarray(batch[0], np.int32), np.asarray(batch[1], np.int32), np.asarray(batch[2], np.int32))
            done += 1
            batch = next(it)
        t_steps = time.perf_counter() - t1
        # charge the sampler its per-step share of a whole epoch
        return t_steps + t_sample * n_steps / self.steps_per_epoch, ref_port.sampler_kind()

    def cpu_tables(self):
        nu = self.d["num_users"]
        return self.E0[:nu], self.E0[nu:]


def make_workload(name, rank):
    return {"bprmf-ml100k": BprmfMl100k, "neumf-ml100k": NeumfMl100k, "lightgcn-gowalla": LightgcnGowalla}[name](rank)


# ----------------------------------------------------------------------------------------
# measurement
# ----------------------------------------------------------------------------------------
def eval_once(w, users):
    from neurec_b200 import ops
    tabs = w.eval_tables()
    if tabs is None:
        return ops.mean_rows(w.run_eval(users))
    res = ops.eval_mf_auto(tabs[0], tabs[1], users, w.tp, w.ti, w.sp, w.si, METRICS, w.eval_k)   # what UniEvaluator calls
    return ops.mean_rows(res)


def measure(w, K, W, world, rank, windows, with_cpu=True):
    """All numbers for one workload.  Returns the dict rank 0 prints (None on other ranks)."""
    import torch
    w.setup_device()
    arrays = w.epoch_arrays(K + W, epoch=0)
    cut = lambda n0: tuple(a[n0 * w.batch:] for a in arrays)

    # ---- value: device-resident steps
    w.run_steps_device(arrays, W)
    barrier(world); flush_l2(); barrier(world)
    wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launches = w.run_steps_device(cut(W), K)
    e1.record()
    barrier(world)
    windows.append((wall0, time.perf_counter()))
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    value = world * K * w.batch / (ms * 1e-3)

    # ---- e2e: per-step host buffers (pinned), loss read back every step
    h_arrays = tuple(a.cpu().pin_memory() for a in arrays)
    w.run_steps_e2e(h_arrays, W)
    barrier(world); flush_l2(); barrier(world)
    wall0 = time.perf_counter()
    _, h2d, d2h = w.run_steps_e2e(tuple(h[W * w.batch:] for h in h_arrays), K)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - wall0, world)
    windows.append((wall0, time.perf_counter()))
    barrier(world)
    e2e_value = world * K * w.batch / e2e_s
    # the same with a host wait after EVERY step (the reference's `sess.run` per batch behaviour)
    ks = min(K, 400)
    barrier(world)
    wall0 = time.perf_counter()
    w.run_steps_e2e_sync(tuple(h[W * w.batch:] for h in h_arrays), ks)
    torch.cuda.synchronize()
    e2e_sync_s = max_over_ranks(time.perf_counter() - wall0, world)
    windows.append((wall0, time.perf_counter()))
    barrier(world)

    # ---- evaluator: users sharded over ranks
    nu = w.d["num_users"]
    n_eval = nu if w.eval_tables() is not None else min(nu, 943)
    from neurec_b200.evaluator import sharded
    a_, b_ = sharded.local_slice(n_eval, rank, world)
    mine = torch.arange(a_, b_, dtype=torch.int32, device="cuda")
    eval_once(w, mine)
    barrier(world); flush_l2(); barrier(world)
    reps = 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        eval_once(w, mine)
    b.record()
    barrier(world)
    windows.append((wall0, time.perf_counter()))
    eval_ms = max_over_ranks(a.elapsed_time(b), world) / reps

    if rank != 0:
        return None
    # ---- roofline: each kernel alone, graph-replayed, CUDA events
    kt, kb, kn = {}, {}, {}
    for name, (fn, nbytes, note) in w.kernels(arrays).items():
        kt[name] = graph_time(fn)
        kb[name], kn[name] = nbytes, note
    dom = max(kt, key=lambda k: kt[k])
    peak, peak_src = measured_peaks()
    achieved = kb[dom] / kt[dom] / 1e9
    out = {
        "metric": "triplets/sec", "value": value, "unit": "triplets/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32",
        "data": "%s (%s), random-init tables, device Philox negatives" % (
            w.d["name"], "the reference's ratio-0.8 split, tests/golden/ml100k_split.npz"
            if w.d["name"] == "ml-100k" else "synthetic, seed 7"),
        "config": {"workload": w.describe, "global_batch": w.batch * world,
                   "optimizer": "adam (TensorFlow-1.12 semantics: dense over every table each step)",
                   "parallelism": "replicas x%d (training does not shard at this size); evaluator: users "
                                  "sharded over ranks" % world,
                   "l2": "flushed (256 MiB write) before each timed region; the K dependent steps then run "
                         "back-to-back as in training (model + optimizer state are L2-sized)"},
        "e2e": {"value": e2e_value, "unit": "triplets/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3 / K,
                "how": "nrc_graph_run_steps: the host stages %d batches from the host arrays into a pinned ring and "
                       "launches one captured burst graph; every step has its own H2D node, kernels and loss "
                       "D2H node, the copies of step s+1 overlap the kernels of step s; the host waits and reads "
                       "the losses once per burst" % w.RING,
                "sync_every_step": {"value": world * ks * w.batch / e2e_sync_s, "unit": "triplets/s",
                                    "ms_per_step": e2e_sync_s * 1e3 / ks, "steps": ks}},
        "gpu_launches": launches,
        "eval": {"metric": "eval users/sec", "value": n_eval / (eval_ms * 1e-3), "unit": "users/s",
                 "users": n_eval,
                 "items": w.d["num_items"], "top_k": w.eval_k, "metrics": 5, "ms": eval_ms,
                 "sharding": "users over %d rank(s)" % world},
        "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic(dom), "peak_source": peak_src,
                     "bytes_per_launch": kb[dom], "bytes_note": kn[dom], "launch_us": kt[dom] * 1e6,
                     "all_kernels": {k: {"us": kt[k] * 1e6, "bytes": kb[k], "GBps": kb[k] / kt[k] / 1e9}
                                     for k in kt},
                     "timing": "each kernel alone: 40 launches in a CUDA graph, CUDA events on the replay "
                               "stream, best of 5"},
    }
    if with_cpu:
        threads = os.cpu_count() or 1
        n_cpu = min(w.steps_per_epoch, 40 if w.name == "lightgcn-gowalla" else 2000)
        dt, skind = w.cpu_reference(n_cpu)
        out["cpu_baseline"] = {"value": n_cpu * w.batch / dt, "unit": "triplets/s", "cores": threads,
                               "kind": "port",
                               "sample": "%d steps of %d: reference sampler/batching (%s random_choice) + numpy/"
                                         "scipy restatement of the TF-1.12 step (TensorFlow is not installable "
                                         "offline); numpy elementwise work is single-threaded" % (n_cpu, w.batch, skind)}
        tabs = w.cpu_tables() if hasattr(w, "cpu_tables") else (None, None)
        dte, n_users_cpu, eimpl, eth = w.cpu_eval(threads, tabs[0], tabs[1],
                                                  max_users=3000 if w.name == "lightgcn-gowalla" else None)
        out["eval"]["cpu"] = {"value": n_users_cpu / dte, "unit": "users/s",
                              "kind": "reference" if eimpl == "reference" else "port", "threads": eth,
                              "sample": "%d users: predict + python mask loop + evaluate.h top-K/metrics, "
                                        "test_batch_size 128" % n_users_cpu}
    return out


def measure_synth_sgd(K, W, world, rank, windows, with_cpu=True):
    """BASELINE config 5 scaled to ONE GPU: BPRMF with plain SGD on synthetic tables far larger
    than L2 (4 M users x 8 M items x d=128 = 6.1 GB), users uniform, positives Zipf(1.05),
    batch 2^20 -- the HBM-bound single-pass kernel (nrc_mf_bpr_sgd_fused)."""
    import torch
    from neurec_b200 import ops
    nu, ni, dim, bs, lr = 4_000_000, 8_000_000, 128, 1 << 20, 0.05
    K = min(K, 24)
    g = torch.Generator(device="cuda").manual_seed(3 + rank)
    U = torch.randn(nu, dim, device="cuda", generator=g) * 0.01
    V = torch.randn(ni, dim, device="cuda", generator=g) * 0.01

    def ids(n):
        u = torch.randint(0, nu, (n,), device="cuda", generator=g, dtype=torch.int32)
        x = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
        a = 1.05   # inverse CDF of the continuous Zipf(a) truncated to [1, ni]
        r = (((ni ** (1 - a) - 1) * x + 1) ** (1 / (1 - a))).clamp(1, ni).long() - 1
        p = ((r * 2654435761) % ni).to(torch.int32)          # scatter the hot ranks over the table
        ng = torch.randint(0, ni, (n,), device="cuda", generator=g, dtype=torch.int32)
        return u, p, ng
    u, p, ng = ids((K + W) * bs)
    loss = torch.zeros(1, device="cuda")
    step = lambda s: ops.mf_bpr_sgd_fused(U, V, u[s * bs:(s + 1) * bs], p[s * bs:(s + 1) * bs],
                                          ng[s * bs:(s + 1) * bs], lr, 0.0, loss)
    for s in range(W):
        step(s)
    barrier(world); flush_l2(); barrier(world)
    wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(W, W + K):
        step(s)
    e1.record()
    barrier(world)
    windows.append((wall0, time.perf_counter()))
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    # e2e: ids of every step come from pinned host memory, the loss goes back every step
    hu, hp, hn = (t.cpu().pin_memory() for t in (u, p, ng))
    du, dp, dn = (torch.empty(bs, dtype=torch.int32, device="cuda") for _ in range(3))
    loss_pin = torch.zeros(1).pin_memory()

    def e2e_step(s):
        sl = slice(s * bs, (s + 1) * bs)
        du.copy_(hu[sl], non_blocking=True); dp.copy_(hp[sl], non_blocking=True); dn.copy_(hn[sl], non_blocking=True)
        loss.zero_()
        ops.mf_bpr_sgd_fused(U, V, du, dp, dn, lr, 0.0, loss)
        loss_pin.copy_(loss, non_blocking=True)
        torch.cuda.synchronize()
    for s in range(W):
        e2e_step(s)
    barrier(world)
    wall0 = time.perf_counter()
    for s in range(W, W + K):
        e2e_step(s)
    e2e_s = max_over_ranks(time.perf_counter() - wall0, world)
    windows.append((wall0, time.perf_counter()))
    barrier(world)
    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    nbytes = bs * (24 * dim + 12)
    kt = ms * 1e-3 / K
    out = {"value": world * K * bs / (ms * 1e-3), "unit": "triplets/s", "steps": K, "ms_per_step": ms / K,
           "e2e": {"value": world * K * bs / e2e_s, "unit": "triplets/s", "h2d_bytes_per_step": 12 * bs,
                   "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3 / K},
           "gpu_launches": K,
           "config": {"workload": "BPRMF synthetic %d users x %d items, dim %d (%.1f GB of tables), learner=gd, "
                                  "batch 2^20, users uniform, positives Zipf(1.05), negatives uniform; single-pass "
                                  "fused step (BASELINE config 5 scaled to one GPU)" % (nu, ni, dim, (nu + ni) * dim * 4 / 1e9),
                      "l2": "tables (6.1 GB) and the per-step id arrays are far larger than L2; every step uses "
                            "fresh ids"},
           "roofline": {"kernel": "mf_bpr_sgd_fused_kernel", "bound": "hbm", "achieved": nbytes / kt / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": nbytes / kt / 1e9 / peak,
                        "traffic": profiled_traffic("mf_bpr_sgd_fused_kernel"),
                        "peak_source": peak_src, "bytes_per_launch": nbytes, "launch_us": kt * 1e6,
                        "bytes_note": "SURVEY.md 8(d): (24*d + 12) B per triplet x 2^20 triplets",
                        "timing": "CUDA events around the K timed launches (one kernel per step)"}}
    if with_cpu:
        from oracle import tf_math
        cn_u, cn_i, cbs = 400_000, 800_000, 1 << 14
        rs = np.random.RandomState(0)
        Uc = (rs.randn(cn_u, dim) * 0.01).astype(np.float32); Vc = (rs.randn(cn_i, dim) * 0.01).astype(np.float32)
        cu, cp, cn = rs.randint(0, cn_u, cbs), rs.randint(0, cn_i, cbs), rs.randint(0, cn_i, cbs)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):   # sparse gd step: gather, dot, g, scatter-sub (TF scatter_sub on IndexedSlices)
            pu, qi, qj = Uc[cu], Vc[cp], Vc[cn]
            x = (pu * qi).sum(1) - (pu * qj).sum(1)
            _, gg = tf_math.pairwise_loss_and_grad("bpr", x)
            gg = gg[:, None]
            np.subtract.at(Uc, cu, np.float32(lr) * gg * (qi - qj))
            np.subtract.at(Vc, cp, np.float32(lr) * gg * pu)
            np.subtract.at(Vc, cn, np.float32(lr) * -gg * pu)
        dt = (time.perf_counter() - t0) / reps
        out["cpu_baseline"] = {"value": cbs / dt, "unit": "triplets/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": "numpy restatement of the TF gd step on a host-RAM-sized slice (%d x %d "
                                         "rows, d=%d, batch 2^14), per-triplet cost extrapolates linearly" % (cn_u, cn_i, dim)}
    del U, V
    torch.cuda.empty_cache()
    return out


def measure_synth_eval(K, W, world, rank, windows, with_cpu=True):
    """BASELINE config 4 on ONE GPU per rank: the full-catalogue evaluator over 10 M items x d=128
    with the score step on the tensor cores (nrc_eval_mf_tc: bf16 tcgen05 candidate pass, exact fp32
    re-score, top-20 in the reference's order, Precision/Recall/MAP/NDCG/MRR).  A step = one batch of
    37 888 users (2 waves of 148 CTAs x 128 users); users are sharded over ranks, the item table is
    replicated, no collective in the data path."""
    import torch
    from neurec_b200 import ops
    nu, ni, dim, topk, ub = 1_000_000, 10_000_000, 128, 20, 37_888
    K = max(1, min(K, 8))
    W = 3
    g = torch.Generator(device="cuda").manual_seed(5)
    V = torch.randn(ni, dim, device="cuda", generator=g) * 0.1
    U = torch.randn(nu, dim, device="cuda", generator=g) * 0.1

    def csr(deg, seed):
        gg = torch.Generator(device="cuda").manual_seed(seed)
        idx = torch.randint(0, ni - deg, (nu, deg), device="cuda", generator=gg, dtype=torch.int32).sort(1).values
        idx += torch.arange(deg, device="cuda", dtype=torch.int32)      # strictly increasing rows: no duplicates
        return (torch.arange(nu + 1, device="cuda", dtype=torch.int64) * deg), idx.reshape(-1).contiguous()
    tp, ti = csr(50, 6)
    sp, si = csr(10, 7)
    batch = lambda s: (torch.arange(ub, device="cuda", dtype=torch.int64)
                       + (rank * (K + W) + s) * ub).remainder(nu).to(torch.int32)
    step = lambda users: ops.eval_mf_tc(U, V, users, tp, ti, sp, si, METRICS, topk)
    for s in range(W):
        step(batch(s))
    torch.cuda.synchronize()
    barrier(world)
    wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, kernel_flops, replays = 0.0, 0.0, 0
    batches = [batch(s) for s in range(W, W + K)]
    sums = torch.zeros(len(METRICS) * topk, dtype=torch.float64, device="cuda")
    e0.record()
    for users in batches:
        res = step(users)
        sums += res.sum(0, dtype=torch.float64)
        km, fl = ops.eval_tc_last_launch()     # waits for the candidate kernel only (events on its stream)
        kernel_ms += km; kernel_flops += fl
    if world > 1:                               # the only collective -- metric sums of all ranks
        import torch.distributed as dist
        dist.all_reduce(sums)
    e1.record()
    barrier(world)
    windows.append((wall0, time.perf_counter()))
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    replays = ops.eval_last_undecided()
    # e2e: user ids from pinned host memory, metric rows back to pinned host memory, every step
    h_users = [b.cpu().pin_memory() for b in batches]
    d_users = torch.empty(ub, dtype=torch.int32, device="cuda")
    h_res = torch.empty((ub, len(METRICS) * topk), dtype=torch.float32).pin_memory()

    def e2e_step(h):
        d_users.copy_(h, non_blocking=True)
        r = step(d_users)
        h_res.copy_(r, non_blocking=True)
        torch.cuda.synchronize()
    e2e_step(h_users[0])
    barrier(world)
    wall0 = time.perf_counter()
    for h in h_users:
        e2e_step(h)
    e2e_s = max_over_ranks(time.perf_counter() - wall0, world)
    windows.append((wall0, time.perf_counter()))
    barrier(world)
    mean_ndcg = float(sums.view(len(METRICS), topk)[METRICS.index("NDCG"), topk - 1] / (world * K * ub))
    if rank != 0:
        return None
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tpeak = pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops") or 2250.0
    tsrc = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if pk.get("bf16_tflops_sustained")
            else "nominal dense bf16 (B200_PROFILING.md fallback)")
    ach = kernel_flops / (kernel_ms * 1e-3) / 1e12
    out = {"value": world * K * ub / (ms * 1e-3), "unit": "users/s", "steps": K, "ms_per_step": ms / K,
           "e2e": {"value": world * K * ub / e2e_s, "unit": "users/s", "h2d_bytes_per_step": 4 * ub,
                   "d2h_bytes_per_step": 4 * ub * len(METRICS) * topk, "ms_per_step": e2e_s * 1e3 / K},
           "gpu_launches": 5 * K,
           "config": {"workload": "full-catalogue evaluator, synthetic %d users x %d items, dim %d, top-%d, 5 metrics, "
                                  "train rows of 50 / test rows of 10 items; %d users per step (BASELINE config 4, "
                                  "item table replicated per GPU, users sharded)" % (nu, ni, dim, topk, ub),
                      "l2": "the bf16 item table (2.56 GB) streamed by every step is far larger than L2",
                      "heap_replays_last_step": replays, "ndcg_at_20_all_ranks": mean_ndcg,
                      "collectives": "one all-reduce of the 100 fp64 metric sums at the end (none in the data path)"},
           "roofline": {"kernel": "tc_candidate_kernel", "bound": "tensor", "achieved": ach, "peak": tpeak,
                        "unit": "TFLOP/s", "frac": ach / tpeak, "traffic": profiled_traffic("tc_candidate_kernel"),
                        "peak_source": tsrc, "flops_per_launch": kernel_flops / K, "launch_us": kernel_ms * 1e3 / K,
                        "flops_note": "SURVEY.md 8(d): 2*d flop per (user, item) pair x 37 888 users x 10 M items",
                        "kernel_share_of_step": kernel_ms / ms,
                        "timing": "CUDA events on the launching stream around every timed tc_candidate_kernel "
                                  "launch (nrc_eval_tc_last_launch)"}}
    if with_cpu:
        import oracle
        threads = os.cpu_count() or 1
        n_cpu = 32
        Vh = V.cpu().numpy()
        cu = batches[-1][:256].cpu().numpy().astype(np.int64)
        tph, tih, sph, sih = (x.cpu().numpy() for x in (tp, ti, sp, si))

        def run_cpu(n):
            us = cu[:n]
            Uh = U[torch.from_numpy(us).cuda()].cpu().numpy()
            rows = lambda ptr, idx: oracle.lists_to_csr([idx[ptr[u]:ptr[u + 1]] for u in us])
            trp, tri = rows(tph, tih)
            tep, tei = rows(sph, sih)
            t0 = time.perf_counter()
            want = oracle.eval_mf(Uh, Vh, np.arange(n, dtype=np.int32), trp, tri, tep, tei,
                                  [int(x) for x in ops._metric_arr(METRICS)], topk,
                                  thread_num=threads)
            return time.perf_counter() - t0, want
        dt, want = run_cpu(n_cpu)
        if dt < 4.0:
            n_cpu = int(min(256, max(n_cpu, n_cpu * 10.0 / max(dt, 1e-3)) // 8 * 8))
            dt, want = run_cpu(n_cpu)
        got = res[:n_cpu].cpu().numpy()
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "users/s", "cores": threads, "kind": "port",
                               "sample": "%d users of the last step against the full 10 M-item catalogue: C "
                                         "restatement of MF.predict + the reference's C++ evaluator (OpenMP/AVX2, %d "
                                         "threads)" % (n_cpu, threads),
                               "bit_identical_to_gpu": bool(np.array_equal(got, want))}
    del U, V
    torch.cuda.empty_cache()
    return out


def measure_sharded_sgd(K, W, world, rank, windows):
    """BASELINE config 5, weak-scaled: BPRMF with plain SGD on tables ROW-SHARDED over the ranks,
    6.25 M users x 12.5 M items x d=128 per GPU (= 50 M x 100 M at 8 GPUs).  Every rank trains
    triplets of its own users (uniform), positives Zipf(1.05) over the GLOBAL catalogue, negatives
    uniform over it; remote item rows are read and RED-updated over NVLink by the one fused kernel
    (nrc_mf_bpr_sgd_sharded) -- no collective in the data path, ranks run asynchronously like the
    in-batch hogwild of the single-GPU kernel."""
    import torch
    import torch.distributed as dist
    from neurec_b200 import ops
    from neurec_b200.util import peer
    nu_l, ni_l, dim, bs, lr = 6_250_000, 12_500_000, 128, 1 << 20, 0.05
    K = min(K, 24)
    ni = ni_l * world
    g = torch.Generator(device="cuda").manual_seed(3 + rank)
    myU = torch.randn(nu_l, dim, device="cuda", generator=g) * 0.01
    myV = torch.randn(ni_l, dim, device="cuda", generator=g) * 0.01
    if world > 1:
        Us, Vs = peer.open_peer_shards(myU), peer.open_peer_shards(myV)
    else:
        Us, Vs = [myU], [myV]

    def ids(n):
        u = torch.randint(0, nu_l, (n,), device="cuda", generator=g, dtype=torch.int32) + rank * nu_l
        x = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
        a = 1.05
        r = (((ni ** (1 - a) - 1) * x + 1) ** (1 / (1 - a))).clamp(1, ni).long() - 1
        p = ((r * 2654435761) % ni).to(torch.int32)
        ng = torch.randint(0, ni, (n,), device="cuda", generator=g, dtype=torch.int32)
        return u, p, ng
    u, p, ng = ids((K + W) * bs)
    loss = torch.zeros(1, device="cuda")
    step = lambda s: ops.mf_bpr_sgd_sharded(Us, Vs, rank, u[s * bs:(s + 1) * bs], p[s * bs:(s + 1) * bs],
                                            ng[s * bs:(s + 1) * bs], lr, 0.0, loss)
    for s in range(W):
        step(s)
    torch.cuda.synchronize()
    barrier(world); flush_l2(); barrier(world)
    wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(W, W + K):
        step(s)
    e1.record()
    barrier(world)
    windows.append((wall0, time.perf_counter()))
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    hu, hp, hn = (t.cpu().pin_memory() for t in (u, p, ng))
    du, dp, dn = (torch.empty(bs, dtype=torch.int32, device="cuda") for _ in range(3))
    loss_pin = torch.zeros(1).pin_memory()

    def e2e_step(s):
        sl = slice(s * bs, (s + 1) * bs)
        du.copy_(hu[sl], non_blocking=True); dp.copy_(hp[sl], non_blocking=True); dn.copy_(hn[sl], non_blocking=True)
        loss.zero_()
        ops.mf_bpr_sgd_sharded(Us, Vs, rank, du, dp, dn, lr, 0.0, loss)
        loss_pin.copy_(loss, non_blocking=True)
        torch.cuda.synchronize()
    for s in range(W):
        e2e_step(s)
    barrier(world)
    wall0 = time.perf_counter()
    for s in range(W, W + K):
        e2e_step(s)
    e2e_s = max_over_ranks(time.perf_counter() - wall0, world)
    windows.append((wall0, time.perf_counter()))
    barrier(world)
    remote = float((torch.div(p[:bs].long(), ni_l, rounding_mode="floor") != rank).double().mean())
    finite = bool(torch.isfinite(loss).item())
    del Us, Vs
    barrier(world)
    if rank != 0:
        return None
    peak, peak_src = measured_peaks()
    nbytes = bs * (24 * dim + 12)
    kt = ms * 1e-3 / K
    return {"value": world * K * bs / (ms * 1e-3), "unit": "triplets/s", "steps": K, "ms_per_step": ms / K,
            "e2e": {"value": world * K * bs / e2e_s, "unit": "triplets/s", "h2d_bytes_per_step": 12 * bs,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3 / K},
            "gpu_launches": K,
            "config": {"workload": "BPRMF, learner=gd, tables row-sharded over %d GPU(s): %d users x %d items x d=%d "
                                   "per GPU (%.1f GB per GPU, %d x %d rows in total), batch 2^20 per GPU and step, users "
                                   "of the own shard, positives Zipf(1.05) and negatives uniform over the global "
                                   "catalogue (BASELINE config 5, weak scaling)" % (
                                       world, nu_l, ni_l, dim, (nu_l + ni_l) * dim * 4 / 1e9, nu_l * world, ni),
                       "exchange": "remote item rows are gathered and RED-updated through CUDA-IPC peer mappings over "
                                   "NVLink inside the fused kernel; no NCCL collective in the data path; ranks are "
                                   "not synchronised between steps",
                       "remote_item_row_fraction": remote, "loss_finite": finite,
                       "l2": "tables (9.6 GB per GPU) and the per-step id arrays are far larger than L2"},
            "roofline": {"kernel": "mf_bpr_sgd_fused_kernel", "bound": "hbm", "achieved": nbytes / kt / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": nbytes / kt / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "bytes_per_launch": nbytes, "launch_us": kt * 1e6,
                         "bytes_note": "per GPU: (24*d + 12) B per triplet x 2^20 triplets; with %d ranks %.0f%% of "
                                       "the item-row bytes cross NVLink instead of local HBM" % (world, 100 * remote),
                         "timing": "CUDA events around the K timed launches of the slowest rank"}}


def run_ours(args):
    import torch
    rank, world, local = dist_setup()
    K, W = args.steps, max(args.warmup, 3)
    clocks = ClockSampler(local)
    clocks.start()
    windows = []
    if args.workload == "bprmf-synth":
        o = measure_synth_sgd(K, W, world, rank, windows)
        out = None
        if rank == 0:
            out = {"metric": "triplets/sec", "n_gpus": world, "warmup": W, "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic, seed 3"}
            out.update(o)
        args.only = True
    elif args.workload == "bprmf-sharded":
        if world > 1 and os.environ.get("NRC_EXPERIMENTAL_SHARDED") != "1":
            # Round-1 status: the kernel, the ABI and the host logic exist and the local-row path is
            # verified, but kernels faulted on the CUDA-IPC peer mappings on the 2-GPU box
            # (tests/mgpu_sharded_check.py stage 1); see DESIGN.md section 5.  Refuse instead of crashing.
            if rank == 0:
                emit(json.dumps({"workload": "bprmf-sharded", "n_gpus": world,
                                 "unavailable": "row-sharded peer-memory path not validated yet "
                                                "(set NRC_EXPERIMENTAL_SHARDED=1 to run it anyway)"}))
            return
        o = measure_sharded_sgd(K, W, world, rank, windows)
        out = None
        if rank == 0:
            out = {"metric": "triplets/sec", "n_gpus": world, "warmup": W, "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic, seed 3"}
            out.update(o)
        args.only = True
    elif args.workload == "eval-synth":
        o = measure_synth_eval(K, W, world, rank, windows)
        out = None
        if rank == 0:
            out = {"metric": "eval users/sec", "n_gpus": world, "warmup": 3, "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "bf16 candidates + f32 exact re-score",
                   "data": "synthetic, seed 5"}
            out.update(o)
        args.only = True
    else:
        out = measure(make_workload(args.workload, rank), K, W, world, rank, windows)
    if world == 1 and not args.only:
        others = {}
        for name in WORKLOADS:
            if name == args.workload:
                continue
            k = min(K, 200) if name == "lightgcn-gowalla" else K
            try:
                o = measure(make_workload(name, rank), k, W, world, rank, windows)
                others[name] = {x: o[x] for x in ("value", "unit", "steps", "ms_per_step", "e2e", "eval",
                                                  "roofline", "cpu_baseline", "config", "gpu_launches")}
            except Exception as ex:  # keep the headline line even if a secondary workload fails
                others[name] = {"error": repr(ex)}
        try:
            others["bprmf-synth-sgd"] = measure_synth_sgd(24, W, world, rank, windows)
        except Exception as ex:
            others["bprmf-synth-sgd"] = {"error": repr(ex)}
        try:
            others["eval-synth"] = measure_synth_eval(4, W, world, rank, windows)
        except Exception as ex:
            others["eval-synth"] = {"error": repr(ex)}
        out["others"] = others
    clocks.stop()
    if rank == 0:
        out["clocks"] = clocks.summary(windows)
        emit(json.dumps(out))
    barrier(world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU path on this box's host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = make_workload(args.workload, 0)
    threads = os.cpu_count() or 1
    # bounded sample made of whole epochs where an epoch is short (the reference draws an epoch's
    # negatives up front, so a partial epoch would overcharge it)
    spe = w.steps_per_epoch
    if w.name == "lightgcn-gowalla":
        n = min(max(args.steps, 10), 60)
    else:
        n = spe * min(max(args.steps // spe, 1), 2)
    w.cpu_reference(min(max(args.warmup, 1), 20))
    dt, skind = w.cpu_reference(n)
    value = n * w.batch / dt
    out = {"impl": "reference", "metric": "triplets/sec", "value": value, "unit": "triplets/s",
           "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": dt * 1e3 / n,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "%s, same init tables as the GPU arm" % w.d["name"],
           "config": {"workload": w.describe, "global_batch": w.batch},
           "cpu_baseline": {"value": value, "unit": "triplets/s", "cores": threads, "kind": "port",
                            "sample": "%d steps: reference sampler/batching (%s random_choice) + numpy/scipy "
                                      "restatement of the TF-1.12 step (TensorFlow 1.12 is not installable "
                                      "offline)" % (n, skind)},
           "e2e": {"value": value, "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(out))


def lift_cpu_thread_limits():
    """torchrun exports OMP_NUM_THREADS=1 to its children; the CPU reference legs must be free to
    use every host core (BLAS/OpenMP pools are resized at run time through threadpoolctl)."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything libraries print (NCCL's version banner,
    warnings) was diverted to stderr by main()."""
    data = (line + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line + "\n"); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # stray prints of native libraries -> stderr
    lift_cpu_thread_limits()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1570)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=WORKLOADS + ("bprmf-synth", "eval-synth", "bprmf-sharded"))
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--only", action="store_true", help="measure only --workload (used under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
