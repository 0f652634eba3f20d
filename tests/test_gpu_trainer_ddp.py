"""Data-parallel training on a B200 box: two ranks of `Trainer` against one rank on the concatenated batch.

The reference wraps the model in DistributedDataParallel (image_model/train_JPDVT.py:231): the constructor broadcasts rank
0's parameters (the ranks seed differently, :115-116) and the backward averages the gradients over the ranks before
`opt.step()` / `update_ema` (:368-372).  `Trainer` does the same with one SUM all-reduce of its flat gradient buffer and
a 1/world scale inside the fused AdamW + EMA kernel.  Asserted here, through real process groups:
  * replicas built under different torch seeds hold identical parameters after construction and after a step;
  * a 2-rank step on two half-batches == a 1-rank step on the whole batch: first moment (= the averaged gradient), second
    moment, parameters, EMA - to rounding (the weight-gradient splits meet through fp32 reduction boxes in no fixed
    order, and Adam's first step is sign-like, so a handful of near-zero gradients may flip: compared in aggregate).
Backend: NCCL with one GPU per rank when the box has two, else gloo with both ranks on cuda:0 (NCCL refuses two ranks on
one device); the trainer code path is the same.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CASE = dict(size=96, depth=2, batch=4, grid=3, wseed=31, seed=311, add_mask=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _draws(case):
    from oracle import cases
    d = cases.training_draws(case)
    return d


def _slice(draws, lo, hi):
    return dict(noise_x=draws["noise_x"][lo:hi], perm=draws["perm"], masks=None if draws["masks"] is None else draws["masks"][lo:hi],
                noise_te=draws["noise_te"][lo:hi])


def _build(seed, dev, load_state):
    from jpdvt_mt_ntnu_b200.models import DiT
    from oracle import cases
    torch.manual_seed(seed)
    m = DiT(input_size=CASE["size"], depth=CASE["depth"], hidden_size=768, patch_size=16, num_heads=12)
    if load_state:
        m.load_state_dict(cases.state_for(CASE))
    return m.to(dev)


def _one_step(trainer, diffusion, x, t, piece, draws, steps=1, graph=False):
    kw = dict(block_size=CASE["size"] // CASE["grid"], patch_size=16, add_mask=CASE["add_mask"], grid_size=CASE["grid"])
    loss = None
    if graph:      # a captured step reads the injected noise from fixed device tensors
        draws = dict(draws, noise_x=draws["noise_x"].to(x.device), noise_te=draws["noise_te"].to(x.device))
    for _ in range(steps):
        diffusion._draws = draws
        loss = trainer.step(x, t, piece, graph=graph, **kw)
    return loss


def _steps(mode):
    # graph: call 1 runs eagerly, call 2 captures and replays, call 3 only replays
    return 3 if mode.endswith("graph") else 2


def _worker(rank, world, port, backend, out_path, mode="end"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from jpdvt_mt_ntnu_b200.trainer import Trainer
    from oracle import cases
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 carries the case's weights, rank 1 a fresh init under another seed: only the broadcast can align them
        model = _build(1000 + rank, dev, load_state=(rank == 0))
        d = create_diffusion("")
        if mode == "peer-mc":
            os.environ["JPDVT_PEER_MULTICAST"] = "1"
        if mode == "peer-thread":
            os.environ["JPDVT_PEER_VARIANT"] = "thread"
        tr = Trainer(model, d, lr=1e-3, weight_decay=0.0, ema_decay=0.999, allreduce=mode.split("-")[0])
        assert (tr.px is not None) == mode.startswith("peer")
        n = tr.total                # peer mode pads the flat buffers to world x slice
        def gather(buf):          # gloo has no CUDA all_gather: stage through the host there
            src = (buf if backend == "nccl" else buf.detach().cpu())[:n] if buf.dim() == 1 and buf.numel() >= n else \
                (buf if backend == "nccl" else buf.detach().cpu())
            outs = [torch.empty_like(src) for _ in range(world)]
            dist.all_gather(outs, src.contiguous())
            return outs

        got = gather(tr.p_flat)
        assert torch.equal(got[0], got[1]), "replicas differ after Trainer construction"
        x, t, piece = cases.training_inputs(CASE)
        draws = _draws(CASE)
        half = CASE["batch"] // world
        lo, hi = rank * half, (rank + 1) * half
        loss = _one_step(tr, d, x[lo:hi].to(dev), t[lo:hi].to(dev), piece.to(dev), _slice(draws, lo, hi), steps=_steps(mode),
                         graph=mode.endswith("graph"))
        tr.check_peers()
        if mode.endswith("graph"):
            assert len(tr._graphs) == 1 and "graph" in next(iter(tr._graphs.values())), "the step was not replayed from a CUDA graph"
        if tr.px is not None:
            # the fp32 state is owned slice by slice; rank 0 alone assembles a checkpoint (one-sided peer reads), as the
            # reference's rank-0-only torch.save does under DDP (train_JPDVT.py:409-417)
            own = slice(tr.px.begin, min(tr.px.end, n))
            other = slice(tr.px.chunk * (1 - rank), min(tr.px.chunk * (2 - rank), n))
            mine_before = tr.m_flat[own].clone()
            ck = tr.checkpoint() if rank == 0 else None
            dist.barrier()
            tr.sync_state()
            assert torch.equal(tr.m_flat[own], mine_before)
            if rank == 0:
                flat_names = [k for k, _ in tr.model.named_parameters() if k in tr._named_views(tr.p_flat)]
                for k in flat_names[:6] + flat_names[-6:]:
                    assert torch.equal(ck["model"][k], tr._named_views(tr.p_flat)[k]), k
                    assert torch.equal(ck["ema"][k], tr._named_views(tr.ema_flat)[k]), k
            assert float(tr.m_flat[other].abs().sum()) > 0
        for buf in (tr.p_flat, tr.ema_flat, tr.m_flat, tr.v_flat):
            got = gather(buf)
            assert torch.equal(got[0], got[1]), "replicas diverged after two data-parallel steps"
        losses = gather(loss.reshape(1))
        if rank == 0:
            torch.save({"p": tr.p_flat[:n].cpu(), "ema": tr.ema_flat[:n].cpu(), "m": tr.m_flat[:n].cpu(), "v": tr.v_flat[:n].cpu(),
                        "loss": torch.cat([l.cpu() for l in losses]).mean(), "step": tr.step_count}, out_path)
    finally:
        dist.destroy_process_group()


MODES = ["end", "peer", "peer-mc", "peer-thread", "peer-graph"]


@pytest.mark.parametrize("mode", MODES)
def test_two_rank_step_equals_one_rank_step_on_the_whole_batch(cuda, tmp_path, mode):
    """`end`: NCCL SUM all-reduce + the full optimizer pass on every rank.  `peer*`: the fused reduce-scatter + AdamW/EMA +
    all-gather kernel over NVLink peer memory (csrc/peer_optim.cu) - bulk async copies (default), multimem instructions,
    per-thread peer loads / stores; `peer-graph`: the whole step, exchange included, replayed from a CUDA graph (step count
    and barrier token read from device memory).  Needs two GPUs with symmetric memory, skipped on a one-GPU box."""
    from conftest import rel_l2
    from jpdvt_mt_ntnu_b200.diffusion import create_diffusion
    from jpdvt_mt_ntnu_b200.trainer import Trainer
    from oracle import cases
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    if mode.startswith("peer") and backend != "nccl":
        pytest.skip("the peer-memory step needs two GPUs on one NVLink domain")
    out = str(tmp_path / "two_rank.pt")
    mp.spawn(_worker, args=(2, _free_port(), backend, out, mode), nprocs=2, join=True)
    two = torch.load(out, map_location="cpu")

    dev = torch.device("cuda", 0)
    model = _build(1000, dev, load_state=True)
    p0 = None
    d = create_diffusion("")
    tr = Trainer(model, d, lr=1e-3, weight_decay=0.0, ema_decay=0.999)
    p0 = tr.p_flat.clone().cpu()
    x, t, piece = cases.training_inputs(CASE)
    loss = _one_step(tr, d, x.to(dev), t.to(dev), piece.to(dev), _draws(CASE), steps=_steps(mode))
    assert two["step"] == tr.step_count == _steps(mode)
    # loss of the second step: mean over the whole batch == mean of the two half-batch means
    assert abs(two["loss"].item() - loss.item()) <= 2e-3 * abs(loss.item())
    # first / second moments after two steps are linear / quadratic in the averaged gradients
    assert rel_l2(two["m"], tr.m_flat.cpu()) < 2e-2
    assert rel_l2(two["v"], tr.v_flat.cpu()) < 4e-2
    # parameters: Adam's early steps are sign-like, so compare the update in aggregate and the parameters tightly
    assert rel_l2(two["p"], tr.p_flat.cpu()) < 2e-3
    assert rel_l2(two["p"] - p0, tr.p_flat.cpu() - p0) < 0.15
    assert rel_l2(two["ema"], tr.ema_flat.cpu()) < 1e-4


def test_load_reference_checkpoint(cuda, tmp_path):
    """weights.load_reference_checkpoint follows the reference's inference loader (inference.py:207-211): the file is the
    trainer's dict (train_JPDVT.py:410-416), keys present in the model are loaded with strict=False, the rest reported;
    `use_ema` picks the EMA weights (what valwhiletrain-style evaluation uses).  The loaded model then produces the same
    latents as a model given the state dict directly."""
    from jpdvt_mt_ntnu_b200.models import DiT
    from jpdvt_mt_ntnu_b200.weights import load_reference_checkpoint
    from oracle import cases
    case = cases.FORWARD_CASES["tiny48"]
    state = cases.state_for(case)
    ema = {k: (v * 0.5 if k != "pos_embed" else v.clone()) for k, v in state.items()}
    extra = dict(state)
    extra["y_embedder.embedding_table.weight"] = torch.zeros(3, 768)         # a key the JPDVT model does not have
    shape_clash = dict(ema)
    shape_clash["time_emb_out2.bias"] = torch.zeros(16)                      # wrong shape: skipped, not loaded
    path = str(tmp_path / "0010000.pt")
    torch.save({"model": extra, "ema": shape_clash, "opt": {"state": {}, "param_groups": []}, "args": None, "train_steps": 10000}, path)

    def fresh():
        return DiT(input_size=case["size"], depth=case["depth"], hidden_size=768, patch_size=16, num_heads=12)

    img, t, x_t = cases.forward_inputs(case)
    want_m = fresh()
    want_m.load_state_dict(state)
    want_m.cuda()
    with torch.no_grad():
        want = want_m(img.cuda(), t.cuda(), x_t.cuda())[1]

    m = fresh()
    rep = load_reference_checkpoint(m, path)
    assert rep["train_steps"] == 10000 and rep["loaded"] == len(state)
    assert rep["skipped"] == ["y_embedder.embedding_table.weight"] and rep["missing"] == []
    m.cuda()
    with torch.no_grad():
        got = m(img.cuda(), t.cuda(), x_t.cuda())[1]
    assert torch.equal(got, want)

    m2 = fresh()
    rep2 = load_reference_checkpoint(m2, path, use_ema=True)
    assert rep2["skipped"] == ["time_emb_out2.bias"] and rep2["missing"] == ["time_emb_out2.bias"]
    assert torch.equal(m2.blocks[0].attn.qkv.weight, ema["blocks.0.attn.qkv.weight"])
    # a bare state dict (no trainer wrapper) loads too
    bare = str(tmp_path / "bare.pt")
    torch.save(state, bare)
    m3 = fresh()
    assert load_reference_checkpoint(m3, bare)["loaded"] == len(state)
    assert all(torch.equal(a, b) for a, b in zip(m3.state_dict().values(), want_m.cpu().state_dict().values()))
