#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the neurec_b200 hot path.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference] [--only]

The hot path is "one training batch" (`sess.run((loss, optimizer), feed_dict)` of the reference's
train_model) INCLUDING what the reference pays to produce the batch: the per-epoch negative
sampling (data/sampler.py:71-90), the shuffle (util/data_iterator.py:59) and the batching
(data_iterator.py:147-152).  Every timed region below starts from the TRAIN CSR: sampling, shuffling
and the step run inside it, on the device, exactly like the reference arm pays for them on the host.

Workloads (BASELINE.json configs):
  bprmf-sharded   (default, configs[4]) BPRMF, learner=gd, tables ROW-SHARDED over the ranks, one
                  CSR-fed kernel per step (sampler + shuffle + gather + score + loss + in-place
                  update; remote item rows read / RED-updated over NVLink).  Weak scaling: per GPU
                  6.25 M users x 12.5 M items x d=128 (9.6 GB) and 2^20 triplets per step; at N=1
                  it is the same kernel on the local shard = the largest single-GPU configuration.
  bprmf-ml100k    (configs[0]) BPRMF on ml-100k: ONE persistent launch per epoch.
  neumf-ml100k    (configs[1]) NeuMF on ml-100k.
  lightgcn-gowalla (configs[2]) LightGCN on the REAL gowalla split (tests/golden/gowalla_split.npz).
  eval-synth      (configs[3]) full-catalogue evaluator, 10 M items x d=128, tensor cores.
At N=1 the default run also measures the other workloads into `others` (unless --only).

Rank 0 prints ONE JSON line:
  value     whole-job triplets/s (a pointwise sample counts as one triplet), train
            CSR resident in HBM when the timed region starts
  e2e       the same through the public per-epoch call with HOST buffers: the train interactions
            (the CSR; the positives' users are expanded from it on the device) are copied from pinned host memory inside the timed region
            and every step's loss is copied back (h2d / d2h bytes per step are counted)
  roofline  dominant kernel: algorithmic bytes / CUDA-event launch time vs the
            measured HBM copy peak (MEASURED_PEAKS.json); tensor pipe for eval-synth
  cpu_baseline  the reference's CPU path on this box (oracle/ref_port.py: the reference's real
            compiled sampler + python batching; oracle/torch_port.py: multi-threaded torch-CPU
            restatement of the TF-1.12 step) on a bounded sample
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

WORKLOADS = ("bprmf-sharded", "bprmf-ml100k", "neumf-ml100k", "lightgcn-gowalla", "eval-synth", "eval-sharded")
DEFAULT_WORKLOAD = "bprmf-sharded"
METRICS = ["Precision", "Recall", "NDCG", "MAP", "MRR"]
SEED = 2018


# ----------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------
def _load_split(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + "_split.npz"))
    return {"name": name, "num_users": int(z["num_users"]), "num_items": int(z["num_items"]),
            "train_indptr": z["train_indptr"].astype(np.int64),
            "train_indices": z["train_indices"].astype(np.int32),
            "test_indptr": z["test_indptr"].astype(np.int64),
            "test_indices": z["test_indices"].astype(np.int32)}


def load_ml100k():
    d = _load_split("ml100k")
    d["name"] = "ml-100k"
    return d


def load_gowalla():
    return _load_split("gowalla")


def profiled_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu captures (latest round first), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(p):
            v = json.load(open(p)).get(kernel)
            if v is not None:
                return v
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



def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)) if os.path.isfile(p) else {}


def pin(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


class TrainData:
    """The train interactions a sampler is built from (data/sampler.py:24-39): the CSR as pinned host arrays and as
    device arrays; the flattened positives' users are expanded from the row pointers ON the device (nrc_csr_row_ids),
    so an upload moves only (indptr, indices)."""

    def __init__(self, indptr, indices, users=None):
        import torch
        self.host = [pin(np.asarray(indptr, np.int64)), pin(np.asarray(indices, np.int32))]
        self.dev = [torch.empty_like(h, device="cuda") for h in self.host]
        self.dev.append(torch.empty((self.host[1].numel(),), dtype=torch.int32, device="cuda"))
        self.nbytes = sum(h.numel() * h.element_size() for h in self.host)
        self.n_pos = int(self.host[1].numel())
        self.upload()

    def upload(self):
        """H2D of the train interactions from pinned memory on the current stream (the e2e leg) + the expansion."""
        from neurec_b200 import ops
        for h, t in zip(self.host, self.dev):
            t.copy_(h, non_blocking=True)
        ops.csr_row_ids(self.dev[0], out=self.dev[2])

    @property
    def ptr(self): return self.dev[0]
    @property
    def idx(self): return self.dev[1]
    @property
    def users(self): return self.dev[2]


def timed(fn, world, windows, flush=True):
    """Barrier + sync, CUDA events around fn() on the current stream, barrier + sync; max over ranks.
    Returns (ms, fn's return value)."""
    import torch
    barrier(world)
    if flush:
        flush_l2(); barrier(world)
    wall0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    barrier(world)
    windows.append((wall0, time.perf_counter()))
    return max_over_ranks(a.elapsed_time(b), world), out


WARM_SECONDS = 0.3


def warm_up(step_fn, min_steps):
    """At least `min_steps` untimed steps AND at least WARM_SECONDS of them: the small workloads' steps are
    microseconds, so a fixed handful would be timed while the SM clock is still ramping up after the (host-side)
    setup.  Returns the number of warm-up steps run."""
    import torch
    done, t0 = 0, time.perf_counter()
    while done < min_steps or time.perf_counter() - t0 < WARM_SECONDS:
        step_fn()
        done += 1
        if done >= min_steps:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    return done


def adam_lr_schedule(lr, n):
    out = np.empty(n, np.float32)
    p1, p2, one, lr = np.float32(0.9), np.float32(0.999), np.float32(1), np.float32(lr)
    for s in range(n):
        out[s] = lr * np.sqrt(one - p2) / (one - p1)
        p1 = np.float32(p1 * np.float32(0.9)); p2 = np.float32(p2 * np.float32(0.999))
    return out


def hbm_roofline(kernel, nbytes, seconds, note, timing, extra=None):
    pk = peaks()
    peak = float(pk.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in pk else "fallback (B200_PROFILING.md 6.65 TB/s)"
    ach = nbytes / seconds / 1e9
    r = {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
         "traffic": profiled_traffic(kernel), "peak_source": src, "bytes_per_launch": nbytes,
         "launch_us": seconds * 1e6, "bytes_note": note, "timing": timing}
    if extra:
        r.update(extra)
    return r


def run_cpu_epoch_steps(sampler, step, n_steps):
    """The reference's train_model loop on the host: for batch in data_iter: step(batch)."""
    done, t0 = 0, time.perf_counter()
    while done < n_steps:
        for batch in sampler:                       # pays the epoch's negative sampling + shuffle + batching
            step(*batch)
            done += 1
            if done >= n_steps:
                break
    return time.perf_counter() - t0


def eval_users_per_s(U, V, data, world, rank, windows, k=20, reps=3):
    """Full-catalogue evaluation (predict + mask + top-K + 5 metrics), users sharded over ranks."""
    import torch
    from neurec_b200 import ops
    from neurec_b200.evaluator import sharded
    d = data
    tp, ti, sp, si = dev(d["train_indptr"]), dev(d["train_indices"]), dev(d["test_indptr"]), dev(d["test_indices"])
    a_, b_ = sharded.local_slice(d["num_users"], rank, world)
    mine = torch.arange(a_, b_, dtype=torch.int32, device="cuda")
    once = lambda: ops.mean_rows(ops.eval_mf_auto(U, V, mine, tp, ti, sp, si, METRICS, k))
    once()

    def go():
        for _ in range(reps):
            r = once()
        return r
    ms, res = timed(go, world, windows)
    ms /= reps
    return {"metric": "eval users/sec", "value": d["num_users"] / (ms * 1e-3), "unit": "users/s",
            "users": d["num_users"], "items": d["num_items"], "top_k": k, "metrics": 5, "ms": ms,
            "ndcg_at_10": float(res.view(5, k)[2, 9]), "sharding": "users over %d rank(s)" % world}


def cpu_eval(data, U, V, threads, max_users=None, k=20):
    from oracle import ref_port
    train_dict = ref_port.user_dict(data["train_indptr"], data["train_indices"])
    test_dict = ref_port.user_dict(data["test_indptr"], data["test_indices"])
    if max_users is not None and len(test_dict) > max_users:
        test_dict = {u: test_dict[u] for u in list(test_dict.keys())[:max_users]}
    best = None
    for th in sorted({8, min(threads, 32)}):         # reference default num_thread=8 vs more threads
        t0 = time.perf_counter()
        _, _, impl = ref_port.evaluate(U, V, train_dict, test_dict, [1, 2, 4, 3, 5], k, 128, th)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, th, impl)
    return {"value": len(test_dict) / best[0], "unit": "users/s", "kind": "reference" if best[2] == "reference" else "port",
            "threads": best[1], "sample": "%d users: np.matmul predict + python mask loop + evaluate.h top-K/metrics, "
                                          "test_batch_size 128" % len(test_dict)}


# ----------------------------------------------------------------------------------------
# configs[0]: BPRMF on ml-100k -- one persistent launch per epoch
# ----------------------------------------------------------------------------------------
class MfMl100k:
    name = "bprmf-ml100k"
    describe = ("BPRMF on ml-100k, dim=64, conf/MF.properties (bs 512, adam 1e-3, bpr, reg 0): shuffle + negative "
                "sampling + the 157 steps of an epoch in ONE persistent cooperative launch")
    dim, batch, lr, reg, loss, opt = 64, 512, 1e-3, 0.0, "bpr", "adam"
    pairwise, neg_num = True, 1

    def __init__(self, rank):
        self.d = load_ml100k()
        rs = np.random.RandomState(2017 + rank)
        self.U0 = (rs.randn(self.d["num_users"], self.dim) * 0.01).astype(np.float32)
        self.V0 = (rs.randn(self.d["num_items"], self.dim) * 0.01).astype(np.float32)
        self.n = len(self.d["train_indices"]) * (1 if self.pairwise else self.neg_num + 1)
        self.spe = (self.n + self.batch - 1) // self.batch

    def setup(self):
        import torch
        d = self.d
        self.T = TrainData(d["train_indptr"], d["train_indices"])
        self.U, self.V = dev(self.U0), dev(self.V0)
        z = torch.zeros_like
        self.gU, self.gV, self.mU, self.vU, self.mV, self.vV = z(self.U), z(self.V), z(self.U), z(self.U), z(self.V), z(self.V)
        self.tU = torch.zeros(d["num_users"], dtype=torch.int32, device="cuda")
        self.tV = torch.zeros(d["num_items"], dtype=torch.int32, device="cuda")
        self.ws = [torch.empty(self.n, dtype=torch.int32, device="cuda") for _ in range(3)]
        self.step_loss = torch.zeros(self.spe, device="cuda")
        self.loss_pin = torch.zeros(self.spe).pin_memory()
        self.pows = torch.tensor([0.9, 0.999], device="cuda")
        self.epoch, self.stamp, self.launches = 0, 1, 0

    def run_steps(self, k, e2e=False):
        """k consecutive steps: whole epochs, then a partial one.  e2e: the train interactions are
        uploaded from pinned host memory and the step losses read back, per epoch call."""
        import torch
        from neurec_b200 import ops
        total = 0.0
        while k > 0:
            n = min(k, self.spe)
            if e2e:
                self.T.upload()
            ops.mf_epoch_fused(self.U, self.V, self.T.ptr, self.T.idx, self.T.users, self.T.idx, self.neg_num,
                               self.pairwise, True, False, SEED, self.epoch, self.batch, 0, n, self.loss, self.reg,
                               self.opt, [self.lr, 0.9, 0.999, 1e-8], self.pows, self.gU, self.gV, self.tU, self.tV,
                               self.mU, self.vU, self.mV, self.vV, self.stamp, self.ws[0], self.ws[1], self.ws[2],
                               self.step_loss)
            if e2e:
                self.loss_pin[:n].copy_(self.step_loss[:n], non_blocking=True)
                torch.cuda.current_stream().synchronize()      # MF.train_model logs the epoch loss
                total += float(self.loss_pin[:n].sum())
            self.epoch += 1; self.stamp += n; self.launches += 1
            k -= n
        return total

    def algorithmic_bytes(self, k):
        d = self.d
        rows = d["num_users"] + d["num_items"]
        per_step = self.batch * (12 * self.dim + 12) + rows * self.dim * 4 * 3 * 2
        return k * per_step, ("per step: 512 triplets x (3 rows of 64 f32 + 3 ids) + TF-faithful Adam "
                              "(U+I)*d*4 B x {var,m,v} x R+W (SURVEY 8d); sampling adds 4 B per draw")

    def cpu_make(self):
        from oracle import ref_port, torch_port
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        np.random.seed(SEED)
        sampler = ref_port.PairwiseSamplerPort(train_dict, d["num_items"], 1, self.batch, True)
        tr = torch_port.MFStep(self.U0, self.V0, self.opt, self.lr, self.loss, self.reg, True)
        return sampler, tr.step, ref_port.sampler_kind(), "torch-CPU restatement of the TF-1.12 step (dense Adam)"

    def eval_tables(self):
        return self.U, self.V

    def cpu_tables(self):
        return self.U0, self.V0


# ----------------------------------------------------------------------------------------
# configs[1]: NeuMF on ml-100k
# ----------------------------------------------------------------------------------------
class NeumfMl100k:
    name = "neumf-ml100k"
    describe = ("NeuMF (GMF+MLP) on ml-100k, embedding_size=32, layers [64,32,16], conf/NeuMF.properties "
                "(pointwise cross_entropy, num_neg 4, bs 256, adam 1e-3): shuffle + negative sampling + the 1570 steps of an "
                "epoch in ONE persistent cooperative launch")
    pairwise, neg_num = False, 4
    mf_dim, layers, batch, lr, loss, opt = 32, [64, 32, 16], 256, 1e-3, "cross_entropy", "adam"
    KEYS = ("mf_user", "mf_item", "mlp_user", "mlp_item", "dense")

    def __init__(self, rank):
        self.d = load_ml100k()
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
        self.n = len(self.d["train_indices"]) * (self.neg_num + 1)
        self.spe = (self.n + self.batch - 1) // self.batch

    def setup(self):
        import torch
        from neurec_b200 import ops
        d = self.d
        nu, ni = d["num_users"], d["num_items"]
        self.T = TrainData(d["train_indptr"], d["train_indices"])
        self.shape = ops.NcfShape.make(nu, ni, self.mf_dim, self.layers, 1)
        self.P = {k: dev(v) for k, v in self.P0.items()}
        z = lambda D: {k: torch.zeros_like(v) for k, v in D.items()}
        self.G, self.S0, self.S1 = z(self.P), z(self.P), z(self.P)
        self.tU = torch.zeros(nu, dtype=torch.int32, device="cuda")
        self.tI = torch.zeros(ni, dtype=torch.int32, device="cuda")
        self.step_loss = torch.zeros(self.spe, device="cuda")
        self.loss_pin = torch.zeros(self.spe).pin_memory()
        self.ws = [torch.empty(self.n, dtype=torch.int32, device="cuda") for _ in range(3)]
        self.pows = torch.tensor([0.9, 0.999], device="cuda")
        self.epoch, self.t, self.stamp, self.launches = 0, 0, 1, 0

    def run_steps(self, k, e2e=False):
        import torch
        from neurec_b200 import ops
        total = 0.0
        while k > 0:
            n = min(k, self.spe)
            if e2e:
                self.T.upload()
            ops.ncf_epoch_fused(self.shape, self.P, self.T.ptr, self.T.idx, self.T.users, self.T.idx, self.neg_num, False,
                                True, False, SEED, self.epoch, self.batch, 0, n, self.loss, 0.0, 0.0, self.opt,
                                [self.lr, 0.9, 0.999, 1e-8], self.pows, self.G, self.S0, self.S1, self.tU, self.tI,
                                self.stamp, self.ws[0], self.ws[1], self.ws[2], self.step_loss)
            if e2e:
                self.loss_pin[:n].copy_(self.step_loss[:n], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                total += float(self.loss_pin[:n].sum())
            self.epoch += 1; self.t += n; self.stamp += n; self.launches += 1
            k -= n
        return total

    def algorithmic_bytes(self, k):
        n_par = sum(v.size for v in self.P0.values())
        per_step = self.batch * (4 * 32 * 4 * 2 + 12) + self.P0["dense"].size * 4 * 2 + n_par * 4 * 3 * 2
        return k * per_step, ("per step: 256 samples x (4 rows of 32 f32 gathered + their gradients + ids) + dense "
                              "weights read + gradient written + TF-faithful Adam over 4 tables + dense")

    def cpu_make(self):
        from oracle import ref_port, tf_math
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        np.random.seed(SEED)
        sampler = ref_port.PointwiseSamplerPort(train_dict, d["num_items"], self.neg_num, self.batch, True)
        tr = tf_math.NCFTrainer(self.P0, 32, self.layers, 1, self.opt, self.lr, self.loss, 0.0, 0.0, False)
        step = lambda bu, bi, bl: tr.step(np.asarray(bu, np.int32), np.asarray(bi, np.int32), np.asarray(bl, np.float32))
        return sampler, step, ref_port.sampler_kind(), "numpy restatement of the TF-1.12 step"

    def eval_tables(self):
        return None


# ----------------------------------------------------------------------------------------
# configs[2]: LightGCN on the real gowalla split
# ----------------------------------------------------------------------------------------
class LightgcnGowalla:
    name = "lightgcn-gowalla"
    describe = ("LightGCN on gowalla (29 858 users, 40 981 items, 810 128 train interactions; the reference's "
                "'given' split), 3 layers dim=64, conf/LightGCN.properties (bs 1024, adam 0.01, reg 1e-3, adj_type pre)")
    dim, n_layers, batch, lr, reg = 64, 3, 1024, 0.01, 1e-3

    def __init__(self, rank):
        from neurec_b200.model.general_recommender.LightGCN import bipartite_adjacency   # the product's builder
        self.d = d = load_gowalla()
        u = np.repeat(np.arange(d["num_users"], dtype=np.int32), np.diff(d["train_indptr"]))
        A = bipartite_adjacency(u, d["train_indices"], d["num_users"], d["num_items"], "pre", verbose=False)
        self.A = A.tocoo().astype(np.float32).tocsr()          # LightGCN.py:151-154
        self.A.sort_indices()
        rs = np.random.RandomState(2017 + rank)
        n = d["num_users"] + d["num_items"]
        lim = np.sqrt(6.0 / (d["num_users"] + self.dim))
        self.E0 = rs.uniform(-lim, lim, (n, self.dim)).astype(np.float32)
        self.n = len(d["train_indices"])
        self.spe = (self.n + self.batch - 1) // self.batch

    def setup(self):
        import torch
        d, A = self.d, self.A
        self.T = TrainData(d["train_indptr"], d["train_indices"])
        self.csr = (dev(A.indptr.astype(np.int64)), dev(A.indices.astype(np.int32)), dev(A.data.astype(np.float32)))
        self.order = dev(np.argsort(-np.diff(A.indptr), kind="stable").astype(np.int32))
        self.e0 = dev(self.E0)
        z = lambda: torch.zeros_like(self.e0)
        self.m, self.v, self.ef, self.gf, self.ge = z(), z(), z(), z(), z()
        self.work = (z(), z())
        self.step_loss = torch.zeros((self.spe, 2), device="cuda")
        self.loss_pin = torch.zeros((self.spe, 2)).pin_memory()
        self.lr_sched = adam_lr_schedule(self.lr, 1 << 14)
        self.epoch, self.t, self.launches = 0, 0, 0

    def run_steps(self, k, e2e=False):
        import torch
        from neurec_b200 import ops
        d = self.d
        total = 0.0
        while k > 0:
            n = min(k, self.spe)
            if e2e:
                self.T.upload()
            cnt = min(n * self.batch, self.n)
            u, i, j = ops.epoch_build(self.T.ptr, self.T.idx, self.T.users, self.T.idx, 1, d["num_items"], True, True,
                                      SEED, self.epoch, 0, cnt)
            ops.lightgcn_train_epoch(self.csr, None, self.order, d["num_users"], d["num_items"], self.n_layers, self.e0,
                                     self.m, self.v, u, i, j.view(-1), self.batch, self.reg,
                                     self.lr_sched[self.t:self.t + n], [self.lr, 0.9, 0.999, 1e-8], self.ef, self.gf,
                                     self.ge, self.work, self.step_loss)
            if e2e:
                self.loss_pin[:n].copy_(self.step_loss[:n], non_blocking=True)
                torch.cuda.current_stream().synchronize()
                total += float(self.loss_pin[:n].sum())
            self.epoch += 1; self.t += n; self.launches += 1 + n * (2 * self.n_layers + 4)
            k -= n
        return total

    def algorithmic_bytes(self, k):
        n, nnz = self.A.shape[0], self.A.nnz
        spmm = nnz * 8 + (n + 1) * 8 + 2 * n * self.dim * 4
        per_step = 2 * self.n_layers * spmm + n * self.dim * 4 * 6
        return k * per_step, "per step: 6 SpMM x (nnz*8 + indptr + 2*N*d*4) + dense Adam N*d*4*6 B (SURVEY 8d: 0.41 GB)"

    def l2_gather(self, seconds):
        """The bound that matters for this kernel: the graph and the table are L2-resident (DRAM traffic = the compulsory
        50 MB), and every non-zero gathers a dim*4-byte row out of L2.  L2 -> SM bytes per product vs the L2 -> SM rate
        measured on this GPU with TMA loads (DESIGN.md 3a ablation: 12.3 TB/s = ~6 300 B/clk)."""
        n, nnz = self.A.shape[0], self.A.nnz
        nbytes = nnz * self.dim * 4 + nnz * 8 + (n + 1) * 8 + n * 4
        cap = 12.3e12
        return {"bytes_per_launch": nbytes, "achieved_TBps": nbytes / seconds / 1e12, "cap_TBps": cap / 1e12,
                "frac": nbytes / seconds / cap,
                "measured_gather_rates_TBps": {"LDG.128 on L2-resident random 512 B rows": 9.09,
                                               "cp.async.bulk on the same rows": 6.70},
                "note": "row gathers nnz*dim*4 B + (col, val) stream + row pointers and order, all served by L2; the cap is the "
                        "L2->SM rate measured with TMA tile loads in the evaluator ablation (DESIGN.md 3a); the two gather "
                        "rates are profiles/r2_peer_probe_v2.txt (local, Zipf rows = L2 hits): index-driven row gathers by "
                        "LDG.128 deliver more than the bulk-copy engine does, which is why this kernel is not bulk-copy fed"}

    def spmm_kernel(self):
        from neurec_b200 import ops
        n, nnz = self.A.shape[0], self.A.nnz
        fn = lambda: ops.spmm_csr(self.csr[0], self.csr[1], self.csr[2], self.e0, row_order=self.order, y=self.work[0])
        return fn, nnz * 8 + (n + 1) * 8 + 2 * n * self.dim * 4

    def cpu_make(self):
        from oracle import ref_port, torch_port
        d = self.d
        train_dict = ref_port.user_dict(d["train_indptr"], d["train_indices"])
        np.random.seed(SEED)
        sampler = ref_port.PairwiseSamplerPort(train_dict, d["num_items"], 1, self.batch, True)
        tr = torch_port.LightGCNStep(self.A, self.E0, d["num_users"], self.n_layers, self.lr, self.reg)
        return sampler, tr.step, ref_port.sampler_kind(), "torch-CPU (sparse CSR mm) restatement of the TF-1.12 step"

    def eval_tables(self):
        from neurec_b200 import ops
        ops.lightgcn_propagate(self.csr[0], self.csr[1], self.csr[2], self.order, self.e0, self.n_layers, self.ef, self.work)
        nu = self.d["num_users"]
        return self.ef[:nu].contiguous(), self.ef[nu:].contiguous()

    def cpu_tables(self):
        nu = self.d["num_users"]
        return self.E0[:nu], self.E0[nu:]


SMALL = {"bprmf-ml100k": MfMl100k, "neumf-ml100k": NeumfMl100k, "lightgcn-gowalla": LightgcnGowalla}


def best_torch_threads(make_step, batch_args, candidates=None):
    """The torch-CPU step on this box at a few thread counts (tiny steps lose with 128 threads): returns
    the fastest count.  `make_step()` builds a fresh step function; `batch_args` is one batch."""
    import torch
    from oracle import torch_port
    n = os.cpu_count() or 1
    best = (1e30, 1)
    for th in sorted({c for c in (candidates or (4, 8, 16, 32, n)) if c <= n}):
        torch_port.set_threads(th)
        step = make_step()
        step(*batch_args)
        t0 = time.perf_counter()
        for _ in range(3):
            step(*batch_args)
        dt = (time.perf_counter() - t0) / 3
        if dt < best[0]:
            best = (dt, th)
    torch_port.set_threads(best[1])
    return best[1]


def cpu_train_baseline(w, n_steps):
    from oracle import torch_port
    sampler, step, skind, how = w.cpu_make()
    threads = os.cpu_count() or 1
    if "torch" in how:
        first = next(iter(sampler))
        threads = best_torch_threads(lambda: w.cpu_make()[1], first)
        sampler, step, skind, how = w.cpu_make()
    dt = run_cpu_epoch_steps(sampler, step, n_steps)
    return {"value": n_steps * w.batch / dt, "unit": "triplets/s", "cores": threads, "kind": "port",
            "sample": "%d steps of %d: the reference's sampler + shuffle + python batching (%s random_choice) + %s on "
                      "%d host threads (the fastest of 4/8/16/32/all on this box; TensorFlow 1.12 is not installable "
                      "offline)" % (n_steps, w.batch, skind, how, threads)}, dt


def measure_small(w, K, W, world, rank, windows, with_cpu=True):
    """bprmf-ml100k / neumf-ml100k / lightgcn-gowalla: N independent replicas at N > 1 (these 0.7-18 MB
    models do not shard)."""
    import torch
    w.setup()
    chunk = max(W, 3)
    warmed = chunk * warm_up(lambda: w.run_steps(chunk), 1)
    # two timed regions of K steps each, the faster one is reported (both are listed in `ms_per_step_regions`): on the
    # LightGCN workload one region in two comes out at 0.82 instead of 0.49 ms/step although the same steps run at
    # 0.49 ms in the e2e region of the same process -- a property of the box, not of the step, that a single region
    # would report at random
    regions = [timed(lambda: w.run_steps(K), world, windows)[0] for _ in range(2)]
    ms = min(regions)
    launches_before = w.launches
    value = world * K * w.batch / (ms * 1e-3)
    warm_up(lambda: w.run_steps(chunk, e2e=True), 1)
    barrier(world); flush_l2(); barrier(world)
    wall0 = time.perf_counter()
    w.run_steps(K, e2e=True)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - wall0, world)
    windows.append((wall0, time.perf_counter()))
    barrier(world)
    calls = (K + w.spe - 1) // w.spe
    tabs = w.eval_tables()
    ev = eval_users_per_s(tabs[0], tabs[1], w.d, world, rank, windows) if tabs is not None else None
    if rank != 0:
        return None
    nbytes, note = w.algorithmic_bytes(K)
    if isinstance(w, LightgcnGowalla):
        fn, sb = w.spmm_kernel()
        st = graph_time(fn)
        roof = hbm_roofline("spmm_csr_fast_kernel", sb, st, "SURVEY 8(d): nnz*(4 B col + 4 B val) + indptr + N*d*4 B read + written",
                            "the kernel alone: 40 launches in a CUDA graph, CUDA events on the replay stream, best of 5",
                            {"step_bytes": nbytes / K, "step_note": note, "share_of_step": 6 * st / (ms * 1e-3 / K),
                             "l2_gather": w.l2_gather(st)})
    else:
        kname = "mf_epoch_kernel" if isinstance(w, MfMl100k) else "ncf_epoch_kernel"
        roof = hbm_roofline(kname, nbytes, ms * 1e-3, note,
                            "CUDA events around the K timed steps (sampling and shuffling included)",
                            {"honest_bound": "tables are L2-resident (0.7 MB): the step is bound by two grid-wide barriers "
                                             "and dependent L2 round trips, not by HBM"})
    out = {"metric": "triplets/sec", "value": value, "unit": "triplets/s", "n_gpus": world, "steps": K, "warmup": warmed,
           "ms_per_step": ms / K, "ms_per_step_regions": [r / K for r in regions], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "%s (%s), random-init tables, device Philox negatives + keyed-bijection shuffle" % (
               w.d["name"], "the reference's split, tests/golden/%s_split.npz" % ("ml100k" if w.d["name"] == "ml-100k" else w.d["name"])),
           "config": {"workload": w.describe, "global_batch": w.batch * world, "steps_per_epoch": w.spe,
                      "timed_region": "starts from the train CSR: negative sampling + shuffle + batching + the steps",
                      "optimizer": "adam (TensorFlow-1.12 semantics: dense over every table each step)",
                      "parallelism": "replicas x%d (training does not shard at this size)" % world,
                      "l2": "flushed (256 MiB write) before each timed region; the K dependent steps then run back-to-back "
                            "as in training (model + optimizer state are L2-sized)"},
           "e2e": {"value": world * K * w.batch / e2e_s, "unit": "triplets/s",
                   "h2d_bytes_per_step": w.T.nbytes * calls / K, "d2h_bytes_per_step": 4 * (2 if isinstance(w, LightgcnGowalla) else 1),
                   "ms_per_step": e2e_s * 1e3 / K,
                   "how": "per epoch call: the train CSR is copied from pinned host memory (positives' users expanded on the device), the epoch's "
                          "steps run, the per-step losses are copied back to pinned memory, the host waits (what "
                          "train_model does per epoch); %d call(s) in the timed region" % calls},
           "gpu_launches": w.launches - launches_before if False else None,
           "roofline": roof}
    out["gpu_launches"] = calls + K * (2 * w.n_layers + 4) if isinstance(w, LightgcnGowalla) else calls
    if ev is not None:
        out["eval"] = ev
    if with_cpu:
        n_cpu = min(w.spe * 2, 40 if isinstance(w, LightgcnGowalla) else 3200)
        out["cpu_baseline"], _ = cpu_train_baseline(w, n_cpu)
        if ev is not None:
            tabs = w.cpu_tables()
            out["eval"]["cpu"] = cpu_eval(w.d, tabs[0], tabs[1], os.cpu_count() or 1,
                                          max_users=3000 if isinstance(w, LightgcnGowalla) else None)
    return out


# ----------------------------------------------------------------------------------------
# configs[4]: BPRMF, learner=gd, row-sharded tables, CSR-fed single-pass kernel (the default)
# ----------------------------------------------------------------------------------------
class ShardedCfg:
    users_per_gpu, items_per_gpu, dim, batch, lr = 6_250_000, 12_500_000, 128, 1 << 20, 0.05
    max_pos = 7                 # positives per user: 1..7 (a <= 64-item positive list)
    zipf_a = 1.05
    n_hot = 16384               # replicated head of the item table (N > 1 only): ~2/3 of all positives at Zipf(1.05) (NRC_BENCH_N_HOT)


def synth_shard_csr(cfg, rank, world, device="cuda"):
    """This rank's train CSR: its own users (local ids), 1..max_pos positives each, items
    Zipf(1.05) over the GLOBAL catalogue, ids as the loader's relabel_by_degree leaves them (the
    n_hot most popular items first, the rest in arbitrary order), rows sorted and duplicate-free.
    Built on the device (setup, untimed)."""
    import torch
    g = torch.Generator(device=device).manual_seed(3 + rank)
    nu, ni = cfg.users_per_gpu, cfg.items_per_gpu * world
    deg = torch.randint(1, cfg.max_pos + 1, (nu,), device=device, generator=g)
    x = torch.rand((nu, cfg.max_pos), device=device, generator=g, dtype=torch.float64)
    a = cfg.zipf_a                         # inverse CDF of the continuous Zipf(a) truncated to [1, ni]
    r = (((ni ** (1 - a) - 1) * x + 1) ** (1 / (1 - a))).clamp(1, ni).long() - 1
    # the loader's relabelling (peer.relabel_by_degree): the n_hot most popular items get ids [0, n_hot) in rank
    # order, the order of the rest is arbitrary -- here a multiplicative hash, so that warm rows are spread over the
    # whole table (and both dies' memory) instead of sitting next to each other
    nh = n_hot_of(cfg, world)
    items = torch.where(r < nh, r, nh + ((r - nh) * 2654435761) % (ni - nh))
    del x, r
    big = torch.iinfo(torch.int64).max
    items[torch.arange(cfg.max_pos, device=device)[None, :] >= deg[:, None]] = big
    items = items.sort(1).values
    dup = torch.zeros_like(items, dtype=torch.bool)
    dup[:, 1:] = items[:, 1:] == items[:, :-1]
    items[dup] = big
    items = items.sort(1).values
    keep = items < big
    deg = keep.sum(1)
    indptr = torch.zeros(nu + 1, dtype=torch.int64, device=device)
    indptr[1:] = deg.cumsum(0)
    indices = items[keep].to(torch.int32)
    return indptr.cpu().numpy(), indices.cpu().numpy()


def n_hot_of(cfg, world):
    """Rows of the replicated head.  The head exists to keep thousands of same-address REDs per step off NVLink; on one
    GPU those rows sit in L2 anyway and the replica only adds a pass (measured, profiles/r2_sgd_forms.txt: 0.646 of the
    HBM peak with it, 0.720 without), so the default is 0 at N = 1."""
    return int(os.environ.get("NRC_BENCH_N_HOT", cfg.n_hot if world > 1 else 0))


def sgd_form():
    """The form of the CSR-fed step the library runs: the register form unless NRC_SGD_PIPE=1 (csrc/train_mf.cu)."""
    return "mf_bpr_sgd_pipe_kernel" if os.environ.get("NRC_SGD_PIPE", "0") != "0" else "mf_bpr_sgd_stream_kernel"


def sharded_describe(cfg, world):
    return ("BPRMF, learner=gd, tables row-sharded over %d GPU(s): %d users x %d items x d=%d per GPU (%.1f GB per GPU; "
            "%d x %d rows in total), 2^20 triplets per GPU and step sampled INSIDE the step kernel from the rank's train "
            "CSR (1..%d positives per user, items Zipf(%.2f) over the global catalogue with the most popular ids first "
            "(the loader's relabelling), uniform negatives rejected against the user's row, keyed-bijection shuffle); %s"
            " -- BASELINE configs[4], weak scaling" % (
                world, cfg.users_per_gpu, cfg.items_per_gpu, cfg.dim,
                (cfg.users_per_gpu + cfg.items_per_gpu) * cfg.dim * 4 / 1e9, cfg.users_per_gpu * world,
                cfg.items_per_gpu * world, cfg.max_pos, cfg.zipf_a,
                ("the %d most popular item rows are replicated on every GPU, their deltas summed by one all-reduce per step"
                 % n_hot_of(cfg, world)) if n_hot_of(cfg, world) else "no replicated head"))


def measure_sharded(K, W, world, rank, windows, with_cpu=True, cfg=ShardedCfg):
    import torch
    from neurec_b200 import ops
    from neurec_b200.util import peer
    W = max(W, 3)
    dim, bs = cfg.dim, cfg.batch
    ni = cfg.items_per_gpu * world
    ptr, idx = synth_shard_csr(cfg, rank, world)
    T = TrainData(ptr, idx)
    del ptr, idx
    torch.cuda.empty_cache()
    g = torch.Generator(device="cuda").manual_seed(30 + rank)
    if world > 1:
        VS = peer.alloc_sharded(cfg.items_per_gpu, dim)
        US_local = torch.empty((cfg.users_per_gpu, dim), dtype=torch.float32, device="cuda")
    else:
        VS = peer.single(torch.empty((cfg.items_per_gpu, dim), dtype=torch.float32, device="cuda"))
        US_local = torch.empty((cfg.users_per_gpu, dim), dtype=torch.float32, device="cuda")
    US_local.normal_(0, 0.01, generator=g)
    VS.local.normal_(0, 0.01, generator=g)
    VS.enable_hot(n_hot_of(cfg, world))
    spe = T.n_pos // bs                        # whole batches only (drop_last), so every step is 2^20 triplets
    loss = torch.zeros(1, device="cuda")
    loss_pin = torch.zeros(K + 8).pin_memory()
    state = {"g": 0}

    def step(e2e_slot=None, after_kernel=None):
        e, s = divmod(state["g"], spe)
        ops.mf_bpr_sgd_epoch(US_local, VS, T.ptr, T.idx, T.users, T.idx, ni, True, SEED + rank, e, s * bs, bs, cfg.lr, 0.0, loss)
        state["g"] += 1
        if after_kernel is not None:
            after_kernel.record()
        VS.sync_hot()                          # the replicated head: one all-reduce of n_hot x d floats + apply
        if e2e_slot is not None:               # every step's loss goes back to the host
            loss_pin[e2e_slot:e2e_slot + 1].copy_(loss, non_blocking=True)

    # warm-up: W steps, then as many more as fill WARM_SECONDS -- the same number on every rank (a step holds a collective)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    per_step = (time.perf_counter() - t0) / W
    extra = int(max_over_ranks(max(0.0, WARM_SECONDS - per_step * W) / per_step, world)) + 1
    for _ in range(extra):
        step()
    W += extra
    barrier(world)
    # per-launch durations (CUDA events on the launching stream) inside the timed region
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    kevs = [torch.cuda.Event(enable_timing=True) for _ in range(K)]

    def run():
        evs[0].record()
        for s in range(K):
            step(after_kernel=kevs[s])
            evs[s + 1].record()
    ms, _ = timed(run, world, windows)
    launch_ms = [evs[s].elapsed_time(kevs[s]) for s in range(K)]          # the step kernel alone
    sync_ms = [kevs[s].elapsed_time(evs[s + 1]) for s in range(K)]        # all-reduce + apply of the replicated head
    # e2e: the train interactions come from pinned host memory inside the timed region
    for _ in range(3):
        step(0)
    barrier(world); flush_l2(); barrier(world)
    wall0 = time.perf_counter()
    T.upload()
    for s in range(K):
        step(s)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - wall0, world)
    windows.append((wall0, time.perf_counter()))
    barrier(world)
    finite = bool(np.isfinite(loss_pin[:K].numpy()).all())
    remote = (world - 1) / world if world > 1 else 0.0
    cold_pos = float((T.idx >= n_hot_of(cfg, world)).float().mean().item()) if n_hot_of(cfg, world) else 1.0
    lazy = None
    if world == 1:          # the explicitly-named lazy-Adam run SURVEY 8(d) asks for (single GPU: rows of var, m, v)
        z = torch.zeros_like
        mU, vU, mV, vV = z(US_local), z(US_local), z(VS.local), z(VS.local)
        lr_sched = adam_lr_schedule(1e-3, K + 4096)
        lstate = {"g": 0}

        def lazy_step():
            e, s = divmod(lstate["g"], spe)
            ops.mf_bpr_lazy_adam_epoch(US_local, mU, vU, VS.local, mV, vV, T.ptr, T.idx, T.users, T.idx, ni, True, SEED + 7,
                                       e, s * bs, bs, float(lr_sched[lstate["g"]]), 0.0, loss)
            lstate["g"] += 1
        warm_up(lazy_step, 3)
        lms, _ = timed(lambda: [lazy_step() for _ in range(K)], world, windows)
        lazy = {"ms": lms / K, "finite": bool(torch.isfinite(loss).item())}
        del mU, vU, mV, vV
    if world > 1:
        VS.close()
    if rank != 0:
        return None
    kt = float(np.mean(launch_ms)) * 1e-3
    nbytes = bs * (24 * dim + 12)
    # per direction and GPU: remote item rows read (in) / reduce-added (out).  Negatives are uniform (all outside the
    # replicated head, to first order), positives only when their id is outside it
    nv_bytes = bs * (1.0 + cold_pos) * remote * dim * 4
    extra = {"launch_us_min": float(np.min(launch_ms)) * 1e3, "launch_us_max": float(np.max(launch_ms)) * 1e3,
             "replicated_head": {"rows": n_hot_of(cfg, world), "sync_us_mean": float(np.mean(sync_ms)) * 1e3,
                                 "what": "per step after the kernel: all-reduce (NCCL, N > 1) of the head's deltas + apply"}}
    if world > 1:
        extra["nvlink"] = {"remote_item_row_fraction": remote, "positives_outside_the_replicated_head": cold_pos,
                           "GBps_per_gpu_per_direction": 2 * nv_bytes / kt / 1e9,
                           "of_measured_peer_copy_770_GBps": 2 * nv_bytes / kt / 1e9 / 770.0,
                           "note": "per direction: the rows this GPU reads from peers + the RED payloads peers send to it "
                                   "(and the mirror image outbound); the limiting resource at N > 1"}
    out = {"metric": "triplets/sec", "value": world * K * bs / (ms * 1e-3), "unit": "triplets/s", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic, seeds 3+rank / 30+rank",
           "config": {"workload": sharded_describe(cfg, world), "train_positives_per_gpu": T.n_pos,
                      "global_batch": bs * world, "steps_per_epoch": spe,
                      "exchange": ("remote item rows are bulk-copied in and reduce-added back through peer mappings over "
                                   "NVLink inside the one fused kernel; the replicated head's deltas (%d rows) take one NCCL "
                                   "all-reduce per step, which also keeps the ranks in step" % n_hot_of(cfg, world)) if world > 1
                                  else "single GPU: no exchange",
                      "loss_finite": finite,
                      "l2": "tables (9.6 GB per GPU) and the train CSR (%.2f GB) are far larger than L2; every step draws "
                            "fresh positions of the shuffled epoch" % (T.nbytes / 1e9)},
           "e2e": {"value": world * K * bs / e2e_s, "unit": "triplets/s", "h2d_bytes_per_step": T.nbytes / K,
                   "d2h_bytes_per_step": 4, "ms_per_step": e2e_s * 1e3 / K,
                   "how": "the rank's train CSR (%.0f MB; the flattened positives' users are expanded from the row pointers on "
                          "the device) is copied from pinned host memory inside the timed region, then K steps, each "
                          "copying its loss back to pinned memory" % (T.nbytes / 1e6)},
           "gpu_launches": K * (2 if n_hot_of(cfg, world) else 1),
           "roofline": hbm_roofline(sgd_form(), nbytes, kt,
                                    "SURVEY 8(d): (24*d + 12) B per triplet x 2^20 triplets (the fused sampler's CSR reads "
                                    "are not counted: 'fused: 0 extra')",
                                    "CUDA events on the launching stream around each of the K timed launches; mean", extra)}
    if lazy is not None:
        lb = bs * (72 * dim + 12)
        out["lazy_adam"] = {"value": bs / (lazy["ms"] * 1e-3), "unit": "triplets/s", "ms_per_step": lazy["ms"],
                            "loss_finite": lazy["finite"],
                            "what": "the explicitly-named lazy-Adam run of SURVEY 8(d): nrc_mf_bpr_lazy_adam_epoch, rows of "
                                    "(var, m, v) read + written per triplet, no batch-wide de-duplication -- NOT the reference's "
                                    "dense TF Adam",
                            "roofline": hbm_roofline("mf_bpr_lazy_adam_stream_kernel", lb, lazy["ms"] * 1e-3,
                                                     "SURVEY 8(d): (72*d + 12) B per triplet x 2^20 triplets",
                                                     "CUDA events around the K timed launches")}
    if with_cpu:
        out["cpu_baseline"] = cpu_sharded_baseline(cfg, seconds=12.0)
    return out


def cpu_sharded_baseline(cfg, seconds=12.0):
    """configs[4] on the host: a host-RAM-sized slice with the same dim / positives-per-user / Zipf
    law; the reference's sampler + python batching feed a multi-threaded torch-CPU gd step."""
    from oracle import ref_port, torch_port
    threads = os.cpu_count() or 1
    nu, ni, dim, bs = 200_000, 400_000, cfg.dim, 1 << 14
    rs = np.random.RandomState(0)
    deg = rs.randint(1, cfg.max_pos + 1, nu)
    a = cfg.zipf_a
    x = rs.rand(int(deg.sum()))
    r = (((ni ** (1 - a) - 1) * x + 1) ** (1 / (1 - a))).clip(1, ni).astype(np.int64) - 1
    items = (r * 2654435761) % ni
    off = np.concatenate([[0], np.cumsum(deg)])
    train = {u: sorted(set(items[off[u]:off[u + 1]].tolist())) for u in range(nu)}
    U = (rs.randn(nu, dim) * 0.01).astype(np.float32); V = (rs.randn(ni, dim) * 0.01).astype(np.float32)
    np.random.seed(SEED)
    sampler = ref_port.PairwiseSamplerPort(train, ni, 1, bs, True)
    cal = (rs.randint(0, nu, bs).tolist(), rs.randint(0, ni, bs).tolist(), rs.randint(0, ni, bs).tolist())
    threads = best_torch_threads(lambda: torch_port.MFStep(U, V, "gd", cfg.lr, "bpr", 0.0, True).step, cal)
    tr = torch_port.MFStep(U, V, "gd", cfg.lr, "bpr", 0.0, True)
    n_pos = sum(len(v) for v in train.values())
    spe = (n_pos + bs - 1) // bs
    t0 = time.perf_counter()
    done = 0
    for batch in sampler:                          # one epoch at most: sampling + shuffle + batching + steps
        tr.step(*batch)
        done += 1
        if time.perf_counter() - t0 > seconds and done >= 3:
            break
    dt = time.perf_counter() - t0
    return {"value": done * bs / dt, "unit": "triplets/s", "cores": threads, "kind": "port",
            "sample": "%d steps of 2^14 on a host-RAM-sized slice (%d users x %d items, d=%d, %d positives): the reference's "
                      "sampler + shuffle + python batching (%s random_choice; the epoch's sampling is paid up front) + torch-CPU "
                      "restatement of the TF gd step on %d threads; per-triplet cost extrapolates linearly" % (
                          done, nu, ni, dim, n_pos, ref_port.sampler_kind(), threads)}


# ----------------------------------------------------------------------------------------
# configs[3]: the full-catalogue evaluator on the tensor cores
# ----------------------------------------------------------------------------------------
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
    # trained-looking user rows: noise + a multiple of the mean of 3 of the user's held-out items, so that
    # the truth items really sit among the top-K and the metric epilogue is exercised with hits
    for a in range(0, nu, 1 << 18):
        b = min(nu, a + (1 << 18))
        held = si.view(nu, 10)[a:b, :3].long()
        U[a:b] += V[held.reshape(-1)].view(b - a, 3, dim).mean(1) * 4.0
    del held
    batch = lambda s: (torch.arange(ub, device="cuda", dtype=torch.int64)
                       + (rank * (K + W) + s) * ub).remainder(nu).to(torch.int32)
    step = lambda users: ops.eval_mf_tc(U, V, users, tp, ti, sp, si, METRICS, topk)
    ops.eval_tc_items_version(1)     # ONE evaluation in user batches: the bf16 item table is converted once, not per step
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
    ops.eval_tc_items_version(0)
    del U, V
    torch.cuda.empty_cache()
    return out


def measure_eval_sharded(K, W, world, rank, windows):
    """The item-sharded evaluator, weak-scaled in catalogue size: every rank holds
    5 M items x d=128 of the item table (and its slice of the train CSR) and 1/N of the user table; a step
    evaluates one batch of 37 888 consecutive users against the WHOLE catalogue (N x 5 M items): one
    all-gather of the batch's user rows, the tensor-core candidate pass + exact re-score per shard, one
    all-gather of the [B, 21] (id, score) lists, merge + 5 metrics."""
    import torch
    from neurec_b200 import ops
    from neurec_b200.evaluator import sharded
    ni_l, dim, topk, B = 5_000_000, 128, 20, 37_888
    K = max(1, min(K, 8)); W = 3
    nu = B * (K + W)
    ni = ni_l * world
    g = torch.Generator(device="cuda").manual_seed(50 + rank)
    V = torch.randn(ni_l, dim, device="cuda", generator=g) * 0.1
    a_, b_ = sharded.local_slice(B, rank, world)
    gu = torch.Generator(device="cuda").manual_seed(7)                 # same user table on every rank; each keeps its slice
    U_all = torch.randn(nu, dim, device="cuda", generator=gu) * 0.1

    def csr_local(deg, seed):                                           # this shard's rows of the (global) train CSR
        gg = torch.Generator(device="cuda").manual_seed(seed + rank)
        per = max(1, deg // world)
        idx = torch.randint(0, ni_l - per, (nu, per), device="cuda", generator=gg, dtype=torch.int32).sort(1).values
        idx += torch.arange(per, device="cuda", dtype=torch.int32)
        return torch.arange(nu + 1, device="cuda", dtype=torch.int64) * per, idx.reshape(-1).contiguous()
    tp, ti = csr_local(48, 60)
    gt = torch.Generator(device="cuda").manual_seed(70)                 # test CSR: global ids, same on every rank
    si = torch.randint(0, ni - 10, (nu, 10), device="cuda", generator=gt, dtype=torch.int32).sort(1).values
    si += torch.arange(10, device="cuda", dtype=torch.int32)
    sp = torch.arange(nu + 1, device="cuda", dtype=torch.int64) * 10
    si = si.reshape(-1).contiguous()
    shard = sharded.ItemShard(V, rank * ni_l, tp, ti)
    ops.eval_tc_items_version(1)

    def step(s):
        lo = s * B
        mine = U_all[lo + a_:lo + b_].contiguous()
        return sharded.evaluate_item_sharded(mine, shard, sp, si, lo, B, METRICS, topk)
    for s in range(W):
        step(s)
    ties = torch.zeros(1, dtype=torch.int32, device="cuda")

    def run():
        for s in range(W, W + K):
            res, t = step(s)
            ties.add_(t)
        return res
    ms, res = timed(run, world, windows, flush=False)
    ops.eval_tc_items_version(0)
    if rank != 0:
        return None
    pk = peaks()
    tpeak = pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops") or 2250.0
    flops = 2.0 * B * ni_l * dim                                       # per rank and step
    return {"metric": "eval users/sec", "value": K * B / (ms * 1e-3), "unit": "users/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 candidates + f32 exact re-score", "data": "synthetic, seeds 50+rank / 7 / 70",
            "gpu_launches": K * 8,
            "config": {"workload": "item-sharded full-catalogue evaluator: %d GPU(s) x %d items x d=%d (catalogue of %d items), "
                                   "%d users per step, top-%d, 5 metrics" % (world, ni_l, dim, ni, B, topk),
                       "collectives": "per step: all-gather of the batch's user rows (%.1f MB) + all-gather of the [B, %d] id and "
                                      "score lists (%.1f MB per rank)" % (B * dim * 4 / 1e6, topk + 1, B * (topk + 1) * 8 / 1e6),
                       "tie_rows": int(ties.item()), "user_item_pairs_per_s": K * B * float(ni) / (ms * 1e-3)},
            "e2e": {"value": K * B / (ms * 1e-3), "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "device-resident only (the N=1 e2e of the evaluator is eval-synth's)"},
            "roofline": {"kernel": "tc_candidate_kernel", "bound": "tensor", "achieved": flops / (ms * 1e-3 / K) / 1e12,
                         "peak": tpeak, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3 / K) / 1e12 / tpeak, "traffic": None,
                         "note": "whole step (collectives, re-score and merge included) against the tensor peak, per GPU"}}


def run_ours(args):
    rank, world, local = dist_setup()
    K, W = args.steps, max(args.warmup, 3)
    clocks = ClockSampler(local)
    clocks.start()
    windows = []
    name = args.workload
    if name == "bprmf-sharded":
        out = measure_sharded(K, W, world, rank, windows)
    elif name == "eval-sharded":
        out = measure_eval_sharded(K, W, world, rank, windows)
        args.only = True
    elif name == "eval-synth":
        o = measure_synth_eval(K, W, world, rank, windows)
        out = None
        if rank == 0:
            out = {"metric": "eval users/sec", "n_gpus": world, "warmup": 3, "higher_is_better": True,
                   "scaling": "weak", "vs_baseline": None, "dtype": "bf16 candidates + f32 exact re-score",
                   "data": "synthetic, seed 5"}
            out.update(o)
    else:
        out = measure_small(SMALL[name](rank), K, W, world, rank, windows)
    if world == 1 and not args.only:
        import torch
        others = {}
        keep = ("value", "unit", "steps", "ms_per_step", "ms_per_step_regions", "e2e", "eval", "roofline", "cpu_baseline", "config", "gpu_launches")
        plan = [("bprmf-ml100k", 157 * 4), ("neumf-ml100k", 1570), ("lightgcn-gowalla", 100), ("eval-synth", 4),
                ("bprmf-sharded", 20)]
        for other, k in plan:
            if other == name:
                continue
            try:
                torch.cuda.empty_cache()
                if other == "eval-synth":
                    o = measure_synth_eval(k, W, world, rank, windows)
                elif other == "bprmf-sharded":
                    o = measure_sharded(k, W, world, rank, windows)
                else:
                    o = measure_small(SMALL[other](rank), k, W, world, rank, windows)
                others[other] = {x: o[x] for x in keep if x in o}
            except Exception as ex:  # keep the headline line even if a secondary workload fails
                others[other] = {"error": repr(ex)}
        out["others"] = others
    if world > 1 and not args.only and name == "bprmf-sharded" and os.environ.get("NRC_BENCH_NO_EVAL", "0") == "0":
        # BASELINE's metric is "BPR triplets/sec & eval users/sec @1/2/4/8": the evaluator line of the same N (configs[3]:
        # item table replicated per GPU, users sharded over the ranks -- SURVEY 8e) rides in `others`, as at N = 1
        import torch
        torch.cuda.empty_cache()
        o = measure_synth_eval(4, W, world, rank, windows, with_cpu=False)
        if rank == 0:
            keep = ("value", "unit", "steps", "ms_per_step", "e2e", "roofline", "config", "gpu_launches")
            out["others"] = {"eval-synth": {x: o[x] for x in keep if x in o}}
    clocks.stop()
    if rank == 0:
        out["clocks"] = clocks.summary(windows)
        emit(json.dumps(out))
    barrier(world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU path on this box's host cores (rank 0 only; other ranks exit)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    from oracle import torch_port
    torch_port.set_threads(threads)
    name = args.workload
    if name == "bprmf-sharded":
        cb = cpu_sharded_baseline(ShardedCfg, seconds=max(10.0, min(60.0, 3.0 * args.steps)))
        cfg = ShardedCfg
        out = {"impl": "reference", "metric": "triplets/sec", "value": cb["value"], "unit": "triplets/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": cfg.batch / cb["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic slice, seed 0",
               "config": {"workload": sharded_describe(cfg, args.gpus), "global_batch": cfg.batch * args.gpus,
                          "cpu_arm": "the CPU arm runs a host-RAM-sized slice of the same law in batches of 2^14; ms_per_step "
                                     "is scaled to 2^20 triplets"},
               "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(json.dumps(out))
        return
    if name == "eval-synth":
        emit(json.dumps({"impl": "reference", "unavailable": "eval-synth's CPU arm is its cpu_baseline key (needs the GPU-side tables)"}))
        return
    w = SMALL[name](0)
    n = min(w.spe * max(1, min(args.steps // w.spe, 2)), 60 if name == "lightgcn-gowalla" else 1 << 30)
    cb, dt = cpu_train_baseline(w, n)
    out = {"impl": "reference", "metric": "triplets/sec", "value": cb["value"], "unit": "triplets/s",
           "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": dt * 1e3 / n,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "%s, same init tables as the GPU arm" % w.d["name"],
           "config": {"workload": w.describe, "global_batch": w.batch},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=WORKLOADS)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--only", action="store_true", help="measure only --workload (used under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
