"""B200 training step: `training_losses -> backward -> gradient all-reduce -> AdamW -> EMA`
(image_model/train_JPDVT.py:357-372 with DDP at :231, AdamW at :281, update_ema at :36-46) as one object.

What it changes relative to running the reference loop on the drop-in modules (which also works, see
tests/test_gpu_training.py::test_adamw_step_changes_outputs_and_engines_refresh):
  * parameters, gradients, Adam moments and the EMA copy live in flat fp32 buffers with one shared layout, so the
    optimizer + EMA (+ the bf16 operand refresh) is ONE kernel pass (jpdvt_adamw_ema, 38 B/param) instead of ~450 launches;
  * the gradient all-reduce (NCCL over NVLink, SUM then 1/world inside the optimizer kernel) is issued per backward stage
    on the communicator stream as soon as that stage's gradients exist, overlapping the remaining backward kernels -
    the role DDP's bucketed hooks play in the reference;
  * no per-step host sync (`loss.item()` is left to the caller's logging cadence, train_JPDVT.py:374).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib, parallel, peer
from ._lib import Weights, WeightsT, check, ptr
from .engine import HIDDEN, LATENT
from .training import TrainEngine, _grad_layout, param_grad_map

_BF16_FIELDS = ("w_patch", "w_ada", "w_qkv", "w_proj", "w_fc1", "w_fc2", "w_final", "w_head1")
_STAGE_FIELDS = {
    "head": ("w_final", "b_final", "w_head1", "b_head1", "w_head2", "b_head2"),
    "block": ("w_qkv", "b_qkv", "w_proj", "b_proj", "w_fc1", "b_fc1", "w_fc2", "b_fc2"),
    "embed": ("w_patch", "b_patch", "w_in", "b_in", "t_w0", "t_b0", "t_w2", "t_b2", "w_ada", "b_ada"),
}


class Trainer:
    def __init__(self, model, diffusion, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 ema_decay: float = 0.9999, process_group=None, allreduce: Optional[str] = None):
        _lib.require_device()
        self.lib = _lib.load()
        self.model, self.diffusion = model, diffusion
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        self.group = process_group
        # "stage": all-reduce each backward stage's gradients as soon as they exist (overlaps the rest of the backward, but
        # the NCCL kernels then compete for SMs with the persistent one-CTA-per-SM GEMMs); "end": one all-reduce of the flat
        # gradient buffer after the backward (exposed, ~1.5 ms for 523 MB over NVSwitch, no contention)
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # "peer" (default when the ranks share an NVLink domain): no collective at all - reduce-scatter, AdamW/EMA on the
        # rank's own slice and the all-gather of the bf16 operands are ONE kernel over peer memory (csrc/peer_optim.cu).
        # "end": one NCCL all-reduce of the flat gradient buffer after the backward (exposed, ~1.5 ms for 523 MB), then the
        # full optimizer pass on every rank.  "stage": one all-reduce per backward stage (overlaps the rest of the backward,
        # but the NCCL kernels then compete for SMs with the persistent one-CTA-per-SM GEMMs - measured slower).
        mode = allreduce or os.environ.get("JPDVT_TRAIN_ALLREDUCE")
        auto = mode is None
        if auto:
            mode = "peer" if (self.world > 1 and peer.available(process_group)[0]) else "end"
        if mode not in ("stage", "end", "peer"):
            raise ValueError(f"allreduce must be 'peer', 'end' or 'stage', got {mode!r}")
        if mode == "peer" and self.world > 1:
            ok, why = peer.available(process_group)
            if not ok:
                raise _lib.JpdvtError(f"allreduce='peer' is not possible here: {why}")
        self.allreduce = mode
        self.step_count = 0
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.JpdvtError("Trainer needs the model on a B200 (`.cuda()`)")
        self.device = dev
        model._check_supported()
        d = model.depth
        self.layout = _grad_layout(d)
        sizes = [int(torch.Size(s).numel()) for _, s in self.layout]
        self.total = sum(sizes)
        f32 = dict(device=dev, dtype=torch.float32)
        self.px: Optional[peer.PeerExchange] = None
        if self.allreduce == "peer" and self.world > 1:
            # every flat buffer lives in this rank's symmetric block: the peers read the gradients, write the bf16 operands
            # (and the few fp32 values the kernels read directly) and can read the owner's slice of the fp32 state
            try:
                self.px = peer.PeerExchange(self.total, dev, process_group)
            except Exception as e:  # noqa: BLE001  (no peer access between these GPUs, VMM handles refused, ...)
                if not auto:
                    raise
                self.px, why = None, e
            if auto:
                # the ranks must agree: if the mapping failed anywhere, everybody takes the NCCL path
                ok = torch.tensor([1 if self.px is not None else 0], device=dev, dtype=torch.int32)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
                if int(ok.item()) == 0:
                    if self.px is None and (dist.get_rank(process_group) == 0):
                        import warnings
                        warnings.warn(f"peer-memory optimizer step unavailable ({why}); falling back to the NCCL all-reduce")
                    self.px, self.allreduce = None, "end"
        if self.px is not None:
            self.p_flat, self.m_flat, self.v_flat = self.px.p, self.px.m, self.px.v
            self.pb_flat = self.px.weights_bf16
        else:
            self.p_flat = torch.zeros(self.total, **f32)
            self.m_flat = torch.zeros(self.total, **f32)
            self.v_flat = torch.zeros(self.total, **f32)
            self.pb_flat = torch.zeros(self.total, device=dev, dtype=torch.bfloat16)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int64)     # the step count on the device (graph replays read it)
        self._graphs: Dict = {}
        self.replayed_launches = 0   # library kernels run from graph replays (the host-side launch counter does not see them)
        self._stale = set()          # peer mode: fp32 state buffers whose non-owned slices are behind the owners'
        self.p_views, self.pb_views, self.offsets = {}, {}, {}
        off = 0
        for (name, shape), n in zip(self.layout, sizes):
            self.p_views[name] = self.p_flat[off:off + n].view(shape)
            self.pb_views[name] = self.pb_flat[off:off + n].view(shape)
            self.offsets[name] = (off, n)
            off += n
        # adopt the module's parameters: copy into the flat buffer, then re-point .data at the slices
        by_name = param_grad_map(model, self.p_views)
        with torch.no_grad():
            for name, p in model.named_parameters():
                if name in by_name:
                    by_name[name].copy_(p.data)
                    p.data = by_name[name]
        if self.px is not None:
            self.ema_flat = self.px.ema
            self.ema_flat.copy_(self.p_flat)
        else:
            self.ema_flat = self.p_flat.clone()
        self.pb_flat.copy_(self.p_flat)
        self.pos = model.pos_embed.data[0].contiguous()
        # small derived tensors + transposed copies
        self.b_embed = torch.empty(HIDDEN, **f32)
        self.w_in_t = torch.empty(LATENT, HIDDEN, **f32)
        bf = dict(device=dev, dtype=torch.bfloat16)
        n_mod = d * 6 * HIDDEN + 2 * HIDDEN
        self.wt_t = {
            "w_qkv_t": torch.empty(d, HIDDEN, 3 * HIDDEN, **bf), "w_proj_t": torch.empty(d, HIDDEN, HIDDEN, **bf),
            "w_fc1_t": torch.empty(d, HIDDEN, 4 * HIDDEN, **bf), "w_fc2_t": torch.empty(d, 4 * HIDDEN, HIDDEN, **bf),
            "w_final_t": torch.empty(HIDDEN, HIDDEN, **bf), "w_head1_t": torch.empty(HIDDEN, 64, **bf),
            "w_ada_t": torch.empty(HIDDEN, n_mod, **bf), "t_w2_t": torch.empty(HIDDEN, HIDDEN, **bf),
        }
        self.t_w2_bf16 = torch.empty(HIDDEN, HIDDEN, **bf)
        self.engine = TrainEngine(d, model.input_size, dev)
        w = Weights()
        w.depth, w.tokens, w.image_size, w.reserved = d, self.engine.tokens, model.input_size, 0
        pv, bv = self.p_views, self.pb_views
        for field, t in (("w_patch", bv["w_patch"]), ("b_embed", self.b_embed), ("w_in_t", self.w_in_t), ("pos", self.pos),
                         ("t_w0", pv["t_w0"]), ("t_b0", pv["t_b0"]), ("t_w2", pv["t_w2"]), ("t_b2", pv["t_b2"]),
                         ("w_ada", bv["w_ada"]), ("b_ada", pv["b_ada"]), ("w_qkv", bv["w_qkv"]), ("b_qkv", pv["b_qkv"]),
                         ("w_proj", bv["w_proj"]), ("b_proj", pv["b_proj"]), ("w_fc1", bv["w_fc1"]), ("b_fc1", pv["b_fc1"]),
                         ("w_fc2", bv["w_fc2"]), ("b_fc2", pv["b_fc2"]), ("w_final", bv["w_final"]), ("b_final", pv["b_final"]),
                         ("w_head1", bv["w_head1"]), ("b_head1", pv["b_head1"]), ("w_head2", pv["w_head2"]), ("b_head2", pv["b_head2"])):
            setattr(w, field, ptr(t))
        wt = WeightsT()
        for k, v in self.wt_t.items():
            setattr(wt, k, ptr(v))
        self.engine.adopt(w, wt, keepalive=(self,))
        if self.px is not None:
            self.engine.grad_buffer = self.px.grads
            self.px.set_f32_ranges([(off, off + n) for name, (off, n) in self.offsets.items() if name not in _BF16_FIELDS])
            self._install_state_hook()
        model.__dict__["_train_engine"] = self.engine
        model.__dict__["_train_engine_key"] = "adopted"
        model.__dict__["_adopted_by_trainer"] = True
        self._sync_replicas()
        self._refresh_derived()
        self._works: List = []

    # ------------------------------------------------------------------ replica alignment (DDP's constructor broadcast)
    def _sync_replicas(self) -> None:
        """Every rank adopts the group's rank-0 state.  The reference seeds each rank differently before building the model
        (train_JPDVT.py:115-116) and relies on DistributedDataParallel's constructor to broadcast rank 0's parameters
        (:231); the same holds after a checkpoint load (every rank reads the file, rank 0's copy is authoritative)."""
        if self.world <= 1:
            return
        steps = torch.tensor([self.step_count], dtype=torch.int64, device=self.device)
        parallel.broadcast_state((self.p_flat, self.ema_flat, self.m_flat, self.v_flat, steps), self.group)
        self.step_count = int(steps.item())
        self.step_dev.fill_(self.step_count)
        self.pb_flat.copy_(self.p_flat)

    # ------------------------------------------------------------------ derived operand copies
    def _refresh_derived(self) -> None:
        with _lib.on_device(self.device):
            self._refresh_derived_impl()

    def _refresh_derived_impl(self) -> None:
        st = _lib.stream_ptr(self.device)
        torch.add(self.p_views["b_patch"], self.p_views["b_in"], out=self.b_embed)
        self.w_in_t.copy_(self.p_views["w_in"].t())
        self.t_w2_bf16.copy_(self.p_views["t_w2"])
        d = self.model.depth
        bv = self.pb_views
        if os.environ.get("JPDVT_DGRAD_MN", "1")[:1] != "0":
            # the data-gradient GEMMs read the [out, in] weights themselves (MN-major B operand): only the small timestep-MLP
            # matrix, whose operand copy is not part of jpdvt_weights, still gets a transposed bf16 copy
            check(self.lib.jpdvt_transpose_bf16(ptr(self.t_w2_bf16), ptr(self.wt_t["t_w2_t"]), 1, HIDDEN, HIDDEN, st), "jpdvt_transpose_bf16")
            self.model.__dict__["_epoch"] = self.model.__dict__.get("_epoch", 0) + 1
            return
        for name, src, batch, rows, cols in (("w_qkv_t", bv["w_qkv"], d, 3 * HIDDEN, HIDDEN), ("w_proj_t", bv["w_proj"], d, HIDDEN, HIDDEN),
                                            ("w_fc1_t", bv["w_fc1"], d, 4 * HIDDEN, HIDDEN), ("w_fc2_t", bv["w_fc2"], d, HIDDEN, 4 * HIDDEN),
                                            ("w_final_t", bv["w_final"], 1, HIDDEN, HIDDEN), ("w_head1_t", bv["w_head1"], 1, 64, HIDDEN),
                                            ("w_ada_t", bv["w_ada"], 1, bv["w_ada"].shape[0], HIDDEN), ("t_w2_t", self.t_w2_bf16, 1, HIDDEN, HIDDEN)):
            check(self.lib.jpdvt_transpose_bf16(ptr(src), ptr(self.wt_t[name]), batch, rows, cols, st), "jpdvt_transpose_bf16")
        self.model.__dict__["_epoch"] = self.model.__dict__.get("_epoch", 0) + 1      # invalidates the inference engine's packed copy

    # ------------------------------------------------------------------ gradient all-reduce, one stage at a time
    def _on_stage(self, stage: str, views: Dict[str, torch.Tensor]) -> None:
        if self.world == 1 or self.allreduce != "stage":
            return
        if stage.startswith("block"):
            i = int(stage[5:])
            tensors = [views[f][i] for f in _STAGE_FIELDS["block"]]
        else:
            tensors = [views[f] for f in _STAGE_FIELDS[stage]]
        try:
            with dist._coalescing_manager(group=self.group, device=self.device, async_ops=True) as cm:
                for t in tensors:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._works.append(cm)
        except (AttributeError, TypeError, RuntimeError):
            for t in tensors:
                self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # ------------------------------------------------------------------ one optimisation step
    def step(self, x: torch.Tensor, t: torch.Tensor, time_emb: torch.Tensor, graph: bool = False, **loss_kwargs) -> torch.Tensor:
        """x [B,3,S,S] in [-1,1], t int64 [B], time_emb [1,G*G,8]  ->  mean loss (device scalar, no host sync).
        graph=True: the whole step - scramble, q_sample, forward, loss, backward, gradient exchange, AdamW + EMA, operand
        refresh (train_JPDVT.py:340-372) - is replayed from a CUDA graph; see `_step_graphed`."""
        if graph:
            return self._step_graphed(x, t, time_emb, **loss_kwargs)
        return self._step_body(x, t, time_emb, None, False, **loss_kwargs)

    def _step_body(self, x, t, time_emb, device_draws, dev_step: bool, **loss_kwargs) -> torch.Tensor:
        model = self.model
        model.__dict__["_stage_hook"] = self._on_stage
        self._works = []
        self.diffusion.__dict__["_device_draws"] = device_draws
        try:
            terms = self.diffusion.training_losses(model, x, t, time_emb, None, **loss_kwargs)
        finally:
            self.diffusion.__dict__["_device_draws"] = None
        loss = terms["loss"].mean()
        loss.backward()
        flat = self.engine.last_flat
        self.step_count += 1
        self.step_dev.add_(1)
        hyper = (1.0 / self.world, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.ema_decay)
        st = _lib.stream_ptr(self.device)
        if self.px is not None:
            # one kernel: sum of every rank's gradients for my slice (peer memory), AdamW + EMA on the slice, bf16 operands
            # (+ the fp32 biases) written to every rank; its two in-kernel barriers are the only synchronisation
            with _lib.on_device(self.device):
                state = (ptr(self.p_flat), ptr(self.m_flat), ptr(self.v_flat), ptr(self.ema_flat))
                if dev_step:
                    check(self.lib.jpdvt_adamw_ema_peer_dev(C.byref(self.px.struct), *state, ptr(self.step_dev), *hyper, st),
                          "jpdvt_adamw_ema_peer_dev")
                else:
                    check(self.lib.jpdvt_adamw_ema_peer(C.byref(self.px.next_epoch()), *state, self.step_count, *hyper, st),
                          "jpdvt_adamw_ema_peer")
            self._stale = {"p", "m", "v", "ema"}
        else:
            if self.world > 1 and self.allreduce == "end":
                parallel.sum_gradients(flat, self.group)
            for wk in self._works:
                wk.wait()
            self._works = []
            with _lib.on_device(self.device):
                state = (ptr(self.p_flat), ptr(flat), ptr(self.m_flat), ptr(self.v_flat), ptr(self.ema_flat), ptr(self.pb_flat), self.total)
                if dev_step:
                    check(self.lib.jpdvt_adamw_ema_dev(*state, ptr(self.step_dev), *hyper, st), "jpdvt_adamw_ema_dev")
                else:
                    check(self.lib.jpdvt_adamw_ema(*state, self.step_count, *hyper, st), "jpdvt_adamw_ema")
        self._refresh_derived()
        for p in model.parameters():
            p.grad = None
        self.engine.last_flat = None
        return loss.detach()

    # ------------------------------------------------------------------ the step from a CUDA graph
    def _step_graphed(self, x, t, time_emb, **kw) -> torch.Tensor:
        """Host-free training step.  Per (batch shape, loss keywords): call 1 runs eagerly (every lazy initialisation of the
        library happens outside a capture), call 2 captures the step on static input buffers and replays it, later calls
        only copy the inputs in and replay.  What changes from step to step reaches the graph through device memory: the
        batch, the timesteps, the host-drawn permutation / mask slots (copied into static buffers before the replay, in
        the reference's RNG order), torch's graph-safe Philox state for the two noise draws, and the step count /
        barrier token of the optimizer kernels (`step_dev`, `epoch_dev`).  Not for the NCCL exchange modes."""
        if self.world > 1 and self.px is None:
            raise _lib.JpdvtError("Trainer.step(graph=True) needs the peer-memory exchange (or one GPU), not an NCCL all-reduce")
        grid = int(kw.get("grid_size", 3))
        add_mask = bool(kw.get("add_mask", False))
        key = (tuple(x.shape), x.dtype, tuple(time_emb.shape), tuple(sorted(kw.items())))
        ent = self._graphs.get(key)
        if ent is None:
            self._graphs[key] = {"warm": True}
            return self._step_body(x, t, time_emb, None, False, **kw)
        B, n = x.shape[0], grid * grid
        if "graph" not in ent:
            dev = self.device
            ent["x"] = torch.empty(x.shape, device=dev, dtype=torch.float32)
            ent["t"] = torch.empty(B, device=dev, dtype=torch.int64)
            ent["piece"] = time_emb.to(device=dev, dtype=torch.float32).clone()
            ent["perm"] = torch.empty(n, device=dev, dtype=torch.int32)
            ent["keep"] = torch.empty(B, n, device=dev, dtype=torch.float32) if add_mask else None
            # host-drawn tables are staged through a ring of pinned slots: a slot is rewritten only after the copy that read it
            # has run, so the host runs up to `ring` steps ahead of the device instead of waiting for every replay
            ent["ring"] = 8
            ent["perm_pin"] = torch.empty(ent["ring"], n, dtype=torch.int32).pin_memory()
            ent["keep_pin"] = torch.empty(ent["ring"], B, n, dtype=torch.float32).pin_memory() if add_mask else None
            ent["copied"] = [None] * ent["ring"]
            ent["slot"] = 0
        slot = ent["slot"]
        ent["slot"] = (slot + 1) % ent["ring"]
        if ent["copied"][slot] is not None:
            ent["copied"][slot].synchronize()
        inj = getattr(self.diffusion, "_draws", None)     # parity tests inject the reference's draws (device noise tensors)
        if inj is not None:
            perm = torch.as_tensor(inj["perm"], dtype=torch.int32)
            keep = inj["masks"].to(torch.float32) if add_mask else None
        else:
            perm, keep = self.diffusion.draw_scramble(B, grid, add_mask)
        ent["perm_pin"][slot].copy_(perm)
        with _lib.on_device(self.device):
            ent["x"].copy_(x, non_blocking=True)
            ent["t"].copy_(t, non_blocking=True)
            ent["perm"].copy_(ent["perm_pin"][slot], non_blocking=True)
            if add_mask:
                ent["keep_pin"][slot].copy_(keep)
                ent["keep"].copy_(ent["keep_pin"][slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            ent["copied"][slot] = ev
            if "graph" not in ent:
                draws = {"perm": ent["perm"], "keep": ent["keep"]}
                count, works = self.step_count, self._works
                g = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    ent["loss"] = self._step_body(ent["x"], ent["t"], ent["piece"], draws, True, **kw)
                ent["graph"], ent["launches"] = g, _lib.launch_count() - n0      # library kernels per replay
                # the capture executed nothing: undo its host-side bookkeeping, the replay below is this call's step
                self.step_count, self._works = count, works
            ent["graph"].replay()
        self.step_count += 1
        self.replayed_launches += ent["launches"]
        if self.px is not None:
            self._stale = {"p", "m", "v", "ema"}
        self.model.__dict__["_epoch"] = self.model.__dict__.get("_epoch", 0) + 1
        return ent["loss"]

    # ------------------------------------------------------------------ peer mode: the fp32 state is owned slice by slice
    def sync_state(self, names=("p", "m", "v", "ema")) -> None:
        """Bring this rank's copy of the named fp32 buffers up to date with their owners (one-sided peer reads; a no-op
        outside peer mode or when nothing changed since the last call).  Called by `state_dict()` of the adopted model and
        by the checkpoint methods, so a rank-0-only `save_checkpoint` works as it does under DDP."""
        if self.px is None:
            return
        need = [n for n in names if n in self._stale]
        if need:
            self.px.pull(need)
            self._stale -= set(need)

    def _install_state_hook(self) -> None:
        import weakref
        ref = weakref.ref(self)

        def hook(module, prefix, keep_vars):      # a plain function: copy.deepcopy(model) must not drag the Trainer along
            tr = ref()
            if tr is not None:
                tr.sync_state(("p",))
        self.model.register_state_dict_pre_hook(hook)

    def check_peers(self) -> None:
        """Raise if an in-kernel barrier of the peer-memory step ever timed out (one 4-byte device->host read)."""
        if self.px is not None:
            self.px.check()

    def allreduce_description(self) -> str:
        if self.world == 1:
            return "none (1 GPU)"
        if self.px is not None:
            return self.px.describe()
        mb = self.total * 4 / 1e6
        if self.allreduce == "stage":
            return f"NCCL SUM per backward stage ({mb:.0f} MB in {self.model.depth + 2} pieces), overlapped with the remaining backward"
        return f"one NCCL SUM of the flat fp32 gradient buffer ({mb:.0f} MB) after the backward, 1/world folded into the optimizer kernel"

    # ------------------------------------------------------------------ checkpoint views (train_JPDVT.py:410-416)
    def ema_state_dict(self) -> Dict[str, torch.Tensor]:
        self.sync_state(("ema",))
        views, off = {}, 0
        for (name, shape) in self.layout:
            n = int(torch.Size(shape).numel())
            views[name] = self.ema_flat[off:off + n].view(shape)
            off += n
        out = {k: v.clone() for k, v in param_grad_map(self.model, views).items()}
        out["pos_embed"] = self.model.pos_embed.data.clone()
        return {k: out[k] for k in self.model.state_dict().keys()}

    def optimizer_state(self) -> dict:
        return {"step": self.step_count, "exp_avg": self.m_flat, "exp_avg_sq": self.v_flat, "lr": self.lr, "betas": self.betas,
                "eps": self.eps, "weight_decay": self.weight_decay}

    def _named_views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Views of a buffer laid out like the flat parameter buffer, keyed by the module's parameter names."""
        views = {name: flat[off:off + n].view(shape) for (name, shape), (off, n) in
                 zip(self.layout, (self.offsets[name] for name, _ in self.layout))}
        return param_grad_map(self.model, views)

    def optimizer_state_dict(self) -> dict:
        """The moments in the layout of `torch.optim.AdamW(model.parameters()).state_dict()` - what the reference stores
        under "opt" and feeds back to `opt.load_state_dict` on resume (train_JPDVT.py:281-284, 413): parameter i of
        `model.parameters()` -> {step, exp_avg, exp_avg_sq}; frozen parameters (pos_embed) are listed but carry no state."""
        self.sync_state(("m", "v"))
        names = [n for n, _ in self.model.named_parameters()]
        m, v = self._named_views(self.m_flat), self._named_views(self.v_flat)
        state = {}
        if self.step_count > 0:
            for i, n in enumerate(names):
                if n in m:
                    state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": m[n].detach().clone(),
                                "exp_avg_sq": v[n].detach().clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": True, "params": list(range(len(names)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: dict) -> None:
        names = [n for n, _ in self.model.named_parameters()]
        m, v = self._named_views(self.m_flat), self._named_views(self.v_flat)
        self.m_flat.zero_()
        self.v_flat.zero_()
        steps = set()
        for i, st in sd.get("state", {}).items():
            n = names[int(i)]
            if n not in m:
                continue
            m[n].copy_(st["exp_avg"])
            v[n].copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"optimizer state carries different step counts per parameter: {sorted(steps)}")
        self.step_count = steps.pop() if steps else 0
        groups = sd.get("param_groups") or [{}]
        g = groups[0]
        self.lr, self.eps, self.weight_decay = g.get("lr", self.lr), g.get("eps", self.eps), g.get("weight_decay", self.weight_decay)
        self.betas = tuple(g.get("betas", self.betas))

    def checkpoint(self, args=None) -> dict:
        """The reference trainer's checkpoint dict (train_JPDVT.py:410-416): model, ema, opt, args, train_steps."""
        return {"model": {k: v.detach().clone() for k, v in self.model.state_dict().items()}, "ema": self.ema_state_dict(),
                "opt": self.optimizer_state_dict(), "args": args, "train_steps": self.step_count}

    def save_checkpoint(self, path: str, args=None) -> None:
        torch.save(self.checkpoint(args), path)

    def load_checkpoint(self, ckpt, strict: bool = False) -> int:
        """Resume from a checkpoint written by `save_checkpoint` or by the reference trainer (train_JPDVT.py:238-284:
        model with strict=False, ema, opt, train_steps; missing entries are skipped as the reference does).  Returns
        train_steps."""
        if isinstance(ckpt, (str, os.PathLike)):
            ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
        if "model" in ckpt:
            self.model.load_state_dict(ckpt["model"], strict=strict)        # copies into the flat buffer's slices
        if "ema" in ckpt:
            ema = self._named_views(self.ema_flat)
            for k, val in ckpt["ema"].items():
                if k in ema:
                    ema[k].copy_(val)
        else:
            self.ema_flat.copy_(self.p_flat)
        if ckpt.get("opt") is not None:
            self.load_optimizer_state_dict(ckpt["opt"])
        if ckpt.get("train_steps") is not None and ckpt.get("opt") is None:
            self.step_count = int(ckpt["train_steps"])
        self.pb_flat.copy_(self.p_flat)
        self.step_dev.fill_(self.step_count)
        self._graphs.clear()
        self._stale = set()           # every rank has just loaded the whole state
        # (pos_embed: self.pos is a view of the module's frozen buffer, updated in place by load_state_dict)
        self._sync_replicas()
        self._refresh_derived()
        return int(ckpt.get("train_steps") or self.step_count)


class BatchPrefetcher:
    """Pinned host batches -> device batches, copied one step ahead on a copy stream while the current step computes - the
    job `DataLoader(pin_memory=True)` + `x.to(device)` do at the top of the reference loop (train_JPDVT.py:303-311, 351),
    without the copy sitting in front of every step.  Two device slots; a slot is overwritten only after the compute
    stream has passed the work that read it."""

    def __init__(self, batches, device: torch.device):
        self.batches, self.device = batches, device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots: List[Optional[torch.Tensor]] = [None, None]
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]
        self._used = [False, False]

    def _start(self, host: torch.Tensor, k: int) -> None:
        if self.slots[k] is None or self.slots[k].shape != host.shape or self.slots[k].dtype != host.dtype:
            self.slots[k] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        if self._used[k]:
            self.copy_stream.wait_event(self.consumed[k])
        with torch.cuda.stream(self.copy_stream):
            self.slots[k].copy_(host, non_blocking=True)
            self.copied[k].record(self.copy_stream)

    def __iter__(self):
        it = iter(self.batches)
        try:
            first = next(it)
        except StopIteration:
            return
        k = 0
        self._start(first, k)
        pending = True
        while pending:
            try:
                nxt = next(it)
                self._start(nxt, 1 - k)
            except StopIteration:
                pending = False
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self.copied[k])
            yield self.slots[k]
            self.consumed[k].record(torch.cuda.current_stream(self.device))    # everything enqueued for this batch so far
            self._used[k] = True
            k = 1 - k


class LossLog:
    """Per-step losses read back without stalling the step: each device scalar is copied into a pinned ring (non-blocking);
    `values()` synchronises once - the logging cadence of the reference loop (train_JPDVT.py:374-397) instead of its
    per-step `loss.item()`."""

    def __init__(self, capacity: int = 4096):
        self.ring = torch.empty(capacity, dtype=torch.float32).pin_memory()
        self.n = 0

    def push(self, loss: torch.Tensor) -> None:
        self.ring[self.n % self.ring.numel()].copy_(loss.detach().reshape(()), non_blocking=True)
        self.n += 1

    def values(self, device: Optional[torch.device] = None) -> List[float]:
        torch.cuda.current_stream(device).synchronize()
        k = min(self.n, self.ring.numel())
        return self.ring[:k].tolist()
