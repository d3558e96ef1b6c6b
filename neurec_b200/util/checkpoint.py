"""Checkpoint / resume of a model's device state.

The reference never saves a TensorFlow checkpoint; what it has is pickled pre-trained tables read back at
build time (NeuMF.py:107-118).  On the device path an epoch is milliseconds, so a run is worth resuming only if
EVERYTHING that determines the next step comes back: tables, dense weights, optimizer slots (Adam m / v, ...),
gradient accumulators and touched stamps, TF's fp32 beta-power variables (host mirror + device copy), the step
stamp counter and the position of the process-wide sampler stream (data/sampler.py: epoch counter -- the analogue
of the reference's global rand() / np.random state).  ``save`` writes one ``torch.save`` file (CPU tensors);
``load`` copies into the tensors ``build_graph`` allocated, so a resumed run continues bit for bit.
"""
import torch

from ..data import sampler as _sampler

FORMAT = 1


def _walk(obj, prefix=""):
    """name -> tensor for every CUDA/CPU tensor reachable through attributes, lists, tuples and dicts of `obj`
    (one level of plain containers, nested tuples included)."""
    out = {}

    def visit(name, v):
        if isinstance(v, torch.Tensor):
            out[name] = v
        elif isinstance(v, (list, tuple)):
            for i, x in enumerate(v):
                visit("%s.%d" % (name, i), x)
        elif isinstance(v, dict):
            for k, x in v.items():
                if isinstance(k, (str, int)):
                    visit("%s.%s" % (name, k), x)
    for k, v in vars(obj).items():
        if k in ("evaluator", "logger", "dataset", "sess"):
            continue
        visit(prefix + k, v)
    return out


def _stream_position():
    return _sampler._EPOCH_COUNTER.value


def state_dict(model):
    tensors = _walk(model)
    opt = getattr(model, "opt", None)
    meta = {"format": FORMAT, "model": type(model).__name__, "sampler_stream": _stream_position()}
    if opt is not None:
        tensors.update(_walk(opt, "opt."))
        meta["opt"] = {"kind": opt.kind, "p1": float(opt._p1), "p2": float(opt._p2), "stamp": int(opt.stamp)}
    return {"meta": meta, "tensors": {k: v.detach().cpu().clone() for k, v in tensors.items()}}


def save(model, path):
    """Write the model's full training state to `path`."""
    torch.cuda.synchronize()
    torch.save(state_dict(model), path)


def load(model, path):
    """Restore a state written by ``save`` into a model on which ``build_graph()`` has run."""
    import numpy as np
    state = torch.load(path, map_location="cpu")
    meta = state["meta"]
    if meta.get("format") != FORMAT:
        raise ValueError("unknown checkpoint format %r" % meta.get("format"))
    if meta["model"] != type(model).__name__:
        raise ValueError("checkpoint of %s cannot be loaded into %s" % (meta["model"], type(model).__name__))
    opt = getattr(model, "opt", None)
    if opt is not None and "opt" in meta:
        if meta["opt"]["kind"] != opt.kind:
            raise ValueError("checkpoint was written with learner=%s, the model uses %s" % (meta["opt"]["kind"], opt.kind))
        opt._p1, opt._p2 = np.float32(meta["opt"]["p1"]), np.float32(meta["opt"]["p2"])
        opt.stamp = int(meta["opt"]["stamp"])
        if "opt._pows_dev" in state["tensors"] and opt.device_pows() is None:
            raise ValueError("checkpoint carries Adam beta powers, the model has none")
    live = _walk(model)
    if opt is not None:
        live.update(_walk(opt, "opt."))
    for name, saved in state["tensors"].items():
        if name not in live:
            continue                                       # scratch that the live model has not allocated yet
        if tuple(live[name].shape) != tuple(saved.shape) or live[name].dtype != saved.dtype:
            if live[name].numel() >= saved.numel() and live[name].dim() == 1 and saved.dim() == 1:
                live[name][:saved.numel()].copy_(saved)    # workspaces may have grown
                continue
            raise ValueError("'%s': checkpoint %s %s, model %s %s" % (name, tuple(saved.shape), saved.dtype,
                                                                    tuple(live[name].shape), live[name].dtype))
        live[name].copy_(saved)
    _sampler.reseed(int(meta["sampler_stream"]))
    torch.cuda.synchronize()
    return meta
