"""dgcnn.trainval -- mirror of /root/reference/dgcnn/trainval.py:7-129 (same class and method names).

What changes underneath:
  * the reference builds one TF "tower" per GPU inside ONE process and averages gradients on /cpu:0
    (trainval.py:16,26-29,59-69).  Here it is one process per GPU (torchrun); tower i of `flags.GPUS` is run by
    rank  i * world // len(GPUS); with a single process all towers run back to back on the one device.
  * every trainable variable and its gradient live in ONE flat fp32 buffer.  Each tower's backward adds
    grad/len(GPUS) into it (accumulation over micro-steps is a SUM, trainval.py:79); apply_gradient() issues
    a single NCCL all-reduce(sum) of that buffer -- the only collective -- and one fused TF-form Adam kernel.
  * with more than one rank, the LAST micro-step before apply_gradient() (accum_gradient(..., last=True)) reduces the
    buffer itself, as two buckets: the head's gradients (MergedEdgeConv / FC / Final: 96 % of the bytes, finished first)
    are all-reduced on NCCL's stream while the EdgeConv backward is still running, the rest when backward ends
    (parallel.GradBuckets; both collectives are nodes of the captured CUDA graph).  DGCNN_OVERLAP_AR=0 keeps the single
    all-reduce inside apply_gradient().
  * BatchNorm statistics stay per tower micro-batch (the reference's BN is per tower), so no SyncBN.
  * `sess` is accepted and ignored.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _native as nv
from . import model as _model
from .parallel import GradBuckets, allreduce_flat_, head_split_offset, tower_assignment
from .variables import VariableStore, set_default_store, default_store


class trainval(object):
    ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8  # tf.train.AdamOptimizer defaults

    def __init__(self, flags):
        self._flags = flags
        self._store = None

    # ------------------------------------------------------------------ setup
    def _pick_device(self):
        if not torch.cuda.is_available():
            raise RuntimeError("dgcnn.trainval needs a CUDA device (B200); there is no CPU fallback")
        self._world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._rank = dist.get_rank() if self._world > 1 else 0
        local = int(os.environ.get("LOCAL_RANK", "0")) if self._world > 1 else 0
        dev = torch.device("cuda", local % torch.cuda.device_count())
        torch.cuda.set_device(dev)
        return dev

    def initialize(self):
        f = self._flags
        self._device = self._pick_device()
        # towers this rank executes (main_funcs.py:145-152 cuts one slice per entry of flags.GPUS)
        self._towers = tower_assignment(len(f.GPUS), self._world, self._rank)
        seed = int(getattr(f, "SEED", 0))
        if self._world > 1:
            # flags.py resolves SEED=-1 to time() separately in every process: all ranks take rank 0's value
            t = torch.tensor([seed], dtype=torch.int64, device=self._device)
            dist.broadcast(t, src=0)
            seed = int(t.item())
            f.SEED = seed
        self._store = VariableStore(device=self._device, seed=seed if seed >= 0 else 0)
        old = set_default_store(self._store)
        try:
            with self._store.variable_scope("dgcnn"):                  # trainval.py:29
                _model.declare_variables(f, int(f.NUM_CHANNEL), self._device)
        finally:
            set_default_store(old)
        if f.TRAIN:
            self._store.flatten(extra=2)                                # +2: loss, accuracy ride along
            n = self._store.num_trainable
            self._adam_m = torch.zeros(n, dtype=torch.float32, device=self._device)
            self._adam_v = torch.zeros(n, dtype=torch.float32, device=self._device)
            self._adam_t = 0
        self._setup_overlap()
        self._sync_replicas()
        self.last_loss = None
        self.last_accuracy = None

    def _setup_overlap(self):
        """Gradient all-reduce overlapped with backward (module docstring).  Active with > 1 rank (DGCNN_OVERLAP_AR=force:
        also on a 1-rank group, for tests); needs the head's variables to be the tail of the flat buffer."""
        self._buckets, self._ovl, self._reduced, self._head = None, None, False, []
        mode = os.environ.get("DGCNN_OVERLAP_AR", "1")
        if not self._flags.TRAIN or mode == "0" or not (dist.is_available() and dist.is_initialized()):
            return
        if self._world <= 1 and mode != "force":
            return
        st = self._store
        names = st.trainable_names()
        split = head_split_offset(names, [st.vars[n].numel() for n in names])
        if not split:
            return
        self._buckets = GradBuckets(st.flat_grad, split)
        import atexit
        import weakref
        me = weakref.ref(self)
        atexit.register(lambda: me() is not None and me().release_graphs())   # before the process group goes away
        off = 0
        for n in names:
            if off >= split:
                self._head.append(st.vars[n])
                st.vars[n].register_post_accumulate_grad_hook(self._on_head_grad)
            off += st.vars[n].numel()

    def _on_head_grad(self, param):
        """post-accumulate hook of every head variable.  When the last of them has its gradient (MergedEdgeConv's: the
        head's backward is over, the EdgeConv stack's has not started) the head gradients are folded into the flat buffer
        and their all-reduce starts; the autograd engine goes on with the EdgeConv backward meanwhile."""
        ov = self._ovl
        if ov is None or ov["fired"]:
            return
        ov["seen"] += 1
        if ov["seen"] < len(self._head):
            return
        ov["fired"] = True
        from . import ops as _ops
        # On the SIDE stream (the one the head's weight-gradient GEMMs run on, ordered after everything queued on the main
        # stream so far): the main stream is not held up -- it goes straight on to the EdgeConv backward, which keeps
        # overlapping the dW GEMMs, the fold and the collective.  NCCL's stream forks from the side stream here and is
        # joined by GradBuckets.wait() at the end of the micro-step.
        with torch.cuda.stream(ov["stream"]):
            with _ops._SideStream(self._device):
                pairs = [(ov["views"][id(p)], p.grad) for p in self._head if p.grad is not None]
                if pairs:
                    torch._foreach_add_([v for v, _ in pairs], [g for _, g in pairs], alpha=1.0 / ov["G"])
                    _ops._side_keep.extend(g for _, g in pairs)    # read on the side stream: alive until the join
                for p in self._head:
                    p.grad = None                               # folded: the end-of-backward accumulation skips them
                self._buckets.reduce_head_async()

    def _sync_replicas(self):
        """Every rank starts from rank 0's variables (and optimizer state): replicas that differ would apply the averaged
        gradient to diverging weights.  Also called after restore()."""
        if self._world <= 1:
            return
        st = self._store
        if getattr(st, "flat_param", None) is not None:
            dist.broadcast(st.flat_param, src=0)
            dist.broadcast(self._adam_m, src=0)
            dist.broadcast(self._adam_v, src=0)
        else:
            for name in sorted(st.vars):
                dist.broadcast(st.vars[name].data, src=0)

    @property
    def variables(self) -> VariableStore:
        return self._store

    # ------------------------------------------------------------------ helpers
    def _to_dev(self, a, dtype):
        if a is None:
            return None
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
        return t.to(self._device, dtype=dtype, non_blocking=True)

    def feed_dict(self, data, label=None, weight=None):
        """trainval.py:87-95: per-tower inputs.  Returns {tower: (points, labels, weights)} for this rank's towers,
        already on the device (the host->device copy of the reference's feed_dict)."""
        res = {}
        for i in self._towers:
            res[i] = (self._to_dev(data[i], torch.float32),
                      self._to_dev(label[i], torch.int64) if label is not None else None,
                      self._to_dev(weight[i], torch.float32) if weight is not None else None)
        return res

    def _forward(self, points, labels, weights, dropout_mask=None, want_softmax=True):
        old = set_default_store(self._store)
        try:
            with self._store.variable_scope("dgcnn"):
                pred = _model.build(points, self._flags, dropout_mask=dropout_mask)      # trainval.py:38
        finally:
            set_default_store(old)
        softmax = torch.softmax(pred, dim=-1) if want_softmax else None                   # trainval.py:39
        accuracy = loss = None
        if labels is not None:
            # trainval.py:41-52: accuracy, per-point cross-entropy (x weight), mean -- one fused kernel that also
            # leaves d loss / d logits behind
            from . import ops as _ops
            K = pred.shape[-1]
            loss, accuracy = _ops.softmax_xent(pred.reshape(-1, K), labels.reshape(-1),
                                               weights.reshape(-1) if weights is not None else None)
        return softmax, accuracy, loss

    # ------------------------------------------------------------------ reference API
    def make_summary(self, sess, data, label, weight):
        if not self._flags.TRAIN:
            raise NotImplementedError
        with torch.no_grad():
            res = self.inference(sess, data, label, weight)
        return {"accuracy": float(res[-2]), "loss": float(res[-1])}

    def inference(self, sess, data, label=None, weight=None):
        """trainval.py:103-108 -> [softmax per tower..., (accuracy, loss)]."""
        feeds = self.feed_dict(data, label, weight)
        outs, accs, losses = [], [], []
        with torch.no_grad():
            for i in self._towers:
                pts, lab, wgt = feeds[i]
                sm, acc, loss = self._forward(pts, lab, wgt)
                outs.append(sm)
                if lab is not None:
                    accs.append(acc)
                    losses.append(loss)
        ops = [o.cpu().numpy() for o in outs]
        if label is not None:
            ops += [float(torch.stack(accs).mean()), float(torch.stack(losses).mean())]
        return ops

    # ------------------------------------------------------------------ CUDA-graph replay of a micro-step
    # One tower micro-step (forward + backward into the flat gradient buffer) is ~300 kernel launches of a few
    # microseconds each; issued eagerly the host cannot keep up with the device.  After GRAPH_WARMUP eager runs
    # with the same input shapes (lazy workspaces and function attributes are then in place) the micro-step is
    # captured once into a CUDA graph on static input buffers and replayed.  DGCNN_CUDA_GRAPH=0 disables it.
    GRAPH_WARMUP = 2
    GRAPH_MAX = 4      # captured graphs kept alive (each owns its activations); least recently used is dropped.
                       # Ragged data (-np -1: a different N per batch) therefore cannot grow memory without bound.

    def _as_tensor(self, a):
        return a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))

    def _tower_step_eager(self, pts, lab, wgt, G, reduce_now=False):
        _, acc, loss = self._forward(pts, lab, wgt, want_softmax=False)
        # loss / accuracy ride in the two trailing slots of the flat gradient buffer (tower mean: 1/len(GPUS) each)
        self._store.flat_grad[-2:].add_(torch.stack([loss.detach(), acc.detach().to(loss.dtype)]), alpha=1.0 / G)
        # Every variable's .grad is a view into the flat gradient buffer.  Left in place, autograd would add into each of
        # them with its own small kernel (one per variable); instead the views step aside, backward hands over fresh
        # gradient tensors, and ONE multi-tensor kernel adds them, scaled by 1/len(GPUS) (the tower mean of
        # trainval.py:64-69), into the flat buffer.
        st = self._store
        params = [st.vars[n] for n in st.trainable_names()]
        views = [p.grad for p in params]
        for p in params:
            p.grad = None
        from . import ops as _ops
        old_async = _ops._async_dw
        _ops._async_dw = os.environ.get("DGCNN_ASYNC_DW", "1") != "0"    # weight-gradient GEMMs on a side stream
        if reduce_now:
            self._ovl = {"seen": 0, "fired": False, "G": G, "stream": torch.cuda.current_stream(self._device),
                         "views": {id(p): v for p, v in zip(params, views)}}
        fired = False
        try:
            loss.backward()
            pairs = [(v, p.grad) for v, p in zip(views, params) if p.grad is not None]
        finally:
            fired = bool(self._ovl and self._ovl["fired"])
            self._ovl = None
            _ops.join_side_stream(self._device)
            _ops._async_dw = old_async
            for p, v in zip(params, views):
                p.grad = v
        if pairs:
            torch._foreach_add_([v for v, _ in pairs], [g for _, g in pairs], alpha=1.0 / G)
        if reduce_now:
            if not fired:                                      # (a model whose head saw no gradient: nothing overlapped)
                self._buckets.reduce_head_async()
            self._buckets.reduce_tail()
            self._buckets.wait()
        return acc.detach(), loss.detach()

    def _tower_step(self, data_i, label_i, weight_i, G, reduce_now=False):
        """forward + backward of one tower's micro-batch -> (accuracy, loss) 0-d device tensors.  reduce_now: this is the
        last micro-step before apply_gradient() and the two-bucket gradient all-reduce runs inside it."""
        use_graph = os.environ.get("DGCNN_CUDA_GRAPH", "1") != "0" and label_i is not None
        src_p = self._as_tensor(data_i)
        key = (tuple(src_p.shape), weight_i is not None, bool(reduce_now))
        if not hasattr(self, "_graphs"):
            from collections import OrderedDict
            self._graphs, self._graph_seen = OrderedDict(), {}
        ent = self._graphs.get(key) if use_graph else None
        if ent is not None:
            self._graphs.move_to_end(key)
        if ent is None:
            pts = self._to_dev(data_i, torch.float32)
            lab = self._to_dev(label_i, torch.int64)
            wgt = self._to_dev(weight_i, torch.float32)
            seen = self._graph_seen.get(key, 0)
            if not use_graph or seen < self.GRAPH_WARMUP:
                self._graph_seen[key] = seen + 1
                return self._tower_step_eager(pts, lab, wgt, G, reduce_now)
            # capture: static inputs, private memory pool, side stream (torch.cuda.graph does the stream dance)
            ent = {"pts": pts.clone(), "lab": lab.clone(), "wgt": wgt.clone() if wgt is not None else None}
            torch.cuda.synchronize(self._device)
            n0 = nv.launch_count()
            graph = torch.cuda.CUDAGraph()
            # (with collectives inside, other threads -- NCCL's watchdog -- may query events while this one captures)
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if reduce_now else "global"):
                acc, loss = self._tower_step_eager(ent["pts"], ent["lab"], ent["wgt"], G, reduce_now)
            # the captured kernels have the current scratch buffers' addresses baked in: the entry keeps them alive, so
            # that a later, larger request (another N, an eager inference call) that replaces a buffer in the cache
            # cannot hand this graph's scratch to the allocator while the graph can still be replayed
            ent.update(graph=graph, acc=acc, loss=loss, launches=nv.launch_count() - n0,
                       workspaces=nv.workspaces_snapshot())
            self._graphs[key] = ent
            gmax = int(os.environ.get("DGCNN_CUDA_GRAPH_MAX", self.GRAPH_MAX))
            while len(self._graphs) > max(gmax, 1):
                _, old = self._graphs.popitem(last=False)     # least recently used: frees its private memory pool
                del old
            if len(self._graph_seen) > 4096:                   # bounded bookkeeping for never-repeating shapes
                self._graph_seen.clear()
            # a capture records but does not execute: fall through to the replay below
        else:
            pass
        ent["pts"].copy_(src_p, non_blocking=True)
        ent["lab"].copy_(self._as_tensor(label_i), non_blocking=True)
        if ent["wgt"] is not None:
            ent["wgt"].copy_(self._as_tensor(weight_i), non_blocking=True)
        ent["graph"].replay()
        nv.add_replayed_launches(ent["launches"])
        return ent["acc"], ent["loss"]

    def accum_gradient(self, sess, data, label, weight=None, summary=False, sync=True, last=False):
        """trainval.py:110-119 -> [None, accuracy, loss(, summary)].  Adds this micro-batch's tower-averaged
        gradient into the flat accumulator.  sync=False returns 0-d device tensors instead of floats.
        last=True: the caller promises that apply_gradient() comes next; with more than one rank the gradient
        all-reduce then runs inside this call, overlapped with the backward pass (no effect on the result)."""
        if not self._flags.TRAIN:
            raise NotImplementedError
        if self._reduced:
            raise RuntimeError("accum_gradient(last=True) must be followed by apply_gradient()")
        G = float(len(self._flags.GPUS))
        accs, losses = [], []
        reduce_here = bool(last) and self._buckets is not None
        for i in self._towers:
            acc, loss = self._tower_step(data[i], label[i] if label is not None else None,
                                         weight[i] if weight is not None else None, G,
                                         reduce_now=reduce_here and i == self._towers[-1])
            accs.append(acc)
            losses.append(loss)
        if reduce_here:
            if not self._towers:                   # a rank without a tower still takes part in the collectives
                self._buckets.reduce_head_async()
                self._buckets.reduce_tail()
                self._buckets.wait()
            self._reduced = True
        if losses:
            acc = torch.stack(accs).mean()
            loss = torch.stack(losses).mean()
        else:
            acc = loss = torch.zeros((), device=self._device)
        res = [None, acc, loss] if not sync else [None, float(acc), float(loss)]
        if summary:
            res.append({"accuracy": float(acc), "loss": float(loss)})
        return res

    def release_graphs(self):
        """Drop every captured micro-step (and the memory pools they own).  Captured graphs that contain NCCL nodes keep
        their communicator referenced: dist.destroy_process_group() waits for them, so call this first."""
        if getattr(self, "_graphs", None):
            torch.cuda.synchronize(self._device)
            self._graphs.clear()
            torch.cuda.synchronize(self._device)

    def zero_gradients(self, sess=None):
        if not self._flags.TRAIN:
            raise NotImplementedError
        self._store.flat_grad.zero_()
        self._reduced = False
        return [None]

    def apply_gradient(self, sess=None):
        """trainval.py:125-129: ONE all-reduce of the flat gradient buffer, then fused Adam (TF form)."""
        if not self._flags.TRAIN:
            raise NotImplementedError
        st = self._store
        fg = st.flat_grad
        if self._reduced:
            self._reduced = False                                # the last micro-step reduced both buckets already
        else:
            allreduce_flat_(fg)                                  # towers were pre-divided by len(GPUS)
        n = st.num_trainable
        self.last_loss, self.last_accuracy = fg[-2], fg[-1]      # summed over micro-steps, mean over towers
        self._adam_t += 1
        t = self._adam_t
        lr = float(self._flags.LEARNING_RATE)
        lr_t = lr * math.sqrt(1.0 - self.ADAM_B2 ** t) / (1.0 - self.ADAM_B1 ** t)
        nv.check(nv.lib().dgcnn_adam_tf_step(st.flat_param.data_ptr(), fg.data_ptr(), self._adam_m.data_ptr(),
                                             self._adam_v.data_ptr(), n, lr_t, self.ADAM_B1, self.ADAM_B2,
                                             self.ADAM_EPS, 1.0, nv.stream_ptr(self._device)), "adam_tf_step")
        return None

    # ------------------------------------------------------------------ checkpoints (tf.train.Saver stand-in)
    def save(self, prefix: str, global_step: int) -> str:
        path = "%s-%d" % (prefix, int(global_step))
        blob = {"variables": self._store.state_dict(), "iteration": int(global_step)}
        if self._flags.TRAIN:
            blob.update(adam_m=self._adam_m.cpu(), adam_v=self._adam_v.cpu(), adam_t=self._adam_t)
        torch.save(blob, path)
        return path

    def restore(self, path: str) -> None:
        blob = torch.load(path, map_location="cpu")
        self._store.load_state_dict(blob["variables"])
        if self._flags.TRAIN and "adam_m" in blob:
            self._adam_m.copy_(blob["adam_m"])
            self._adam_v.copy_(blob["adam_v"])
            self._adam_t = int(blob["adam_t"])
        self._sync_replicas()
