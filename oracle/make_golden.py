#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.  TEST INFRASTRUCTURE ONLY.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

The reference (image_model/models.py, image_model/diffusion/, image_model/inference.py) is imported from
where it lies; the two third-party packages it needs that are absent from this image (timm, matplotlib)
are provided by the stand-ins under oracle/standins/.  Nothing from the reference is copied into the repo:
only the *outputs* below are committed, together with this script.

Every random input is regenerated from seeds by tests (see `oracle/cases.py`), so the fixtures only hold
reference OUTPUTS (plus tiny inputs for the assignment cases).
"""
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("JPDVT_REFERENCE", "/root/reference/image_model")
sys.path[:0] = [os.path.join(HERE, "standins"), REF, ROOT]

import numpy as np
import torch

import models as ref_models                      # noqa: E402  (the reference)
from diffusion import create_diffusion           # noqa: E402  (the reference)
import inference as ref_inference                # noqa: E402  (the reference; __main__-guarded)
from sklearn.metrics import pairwise_distances   # noqa: E402

from oracle import cases                          # noqa: E402
from oracle import jpdvt_oracle as orc            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_grad_enabled(False)
torch.set_num_threads(os.cpu_count() or 1)


def build_ref(case):
    m = ref_models.DiT(input_size=case["size"], depth=case["depth"], hidden_size=768, patch_size=16, num_heads=12)
    st = orc.seeded_state(m.state_dict(), seed=case["wseed"], std=case.get("wstd", 0.02))
    m.load_state_dict(st)
    m.train()  # the reference scripts run in train() mode; no dropout/BN so it is a no-op (inference.py:213-214)
    return m, st


def golden_static():
    out = {}
    out["pe_8_3"] = ref_models.get_2d_sincos_pos_embed(8, 3)
    out["pe_8_4"] = ref_models.get_2d_sincos_pos_embed(8, 4)
    out["pe_768_12"] = ref_models.get_2d_sincos_pos_embed(768, 12)
    for name, spec in (("full", ""), ("s250", "250"), ("s10", "10"), ("ddim50", "ddim50"), ("sec", "10,15,20")):
        d = create_diffusion(spec)
        out[f"{name}_map"] = np.asarray(d.timestep_map, dtype=np.int64)
        out[f"{name}_betas"] = d.betas
        out[f"{name}_sqrt_ac"] = d.sqrt_alphas_cumprod
        out[f"{name}_sqrt_1mac"] = d.sqrt_one_minus_alphas_cumprod
        out[f"{name}_post_var"] = d.posterior_variance
        out[f"{name}_post_logvar"] = d.posterior_log_variance_clipped
        out[f"{name}_coef1"] = d.posterior_mean_coef1
        out[f"{name}_coef2"] = d.posterior_mean_coef2
    # fresh-init facts: outputs are exactly zero, parameter count
    m = ref_models.DiT_models["JPDVT"](input_size=192)
    out["n_params_192"] = np.asarray(sum(p.numel() for p in m.parameters()))
    out["n_trainable_192"] = np.asarray(sum(p.numel() for p in m.parameters() if p.requires_grad))
    out["state_keys"] = np.asarray(list(m.state_dict().keys()))
    out["state_numel"] = np.asarray([v.numel() for v in m.state_dict().values()])
    img, te = m(torch.randn(1, 3, 192, 192), torch.tensor([5]), torch.randn(1, 144, 8))
    out["fresh_absmax"] = np.asarray([img.abs().max().item(), te.abs().max().item()])
    np.savez_compressed(os.path.join(OUT, "static.npz"), **out)
    print("static.npz", len(out))


def golden_forward():
    for name, case in cases.FORWARD_CASES.items():
        m, st = build_ref(case)
        img, t, x_t = cases.forward_inputs(case)
        taps = {}
        hooks = []
        for i, blk in enumerate(m.blocks):
            hooks.append(blk.register_forward_hook(lambda mod, a, o, i=i: taps.__setitem__(f"block{i}", o.clone())))
        hooks.append(m.final_layer.register_forward_hook(lambda mod, a, o: taps.__setitem__("final", o.clone())))
        hooks.append(m.t_embedder.register_forward_hook(lambda mod, a, o: taps.__setitem__("c", o.clone())))
        out_img, out_te = m(img, t, x_t)
        for h in hooks:
            h.remove()
        out = {"te": out_te.numpy(), "c": taps["c"].numpy()}
        # full tensors only where small; otherwise a strided sample + moments
        out["img_sample"] = out_img.numpy()[:, :, ::7, ::5]
        out["img_moments"] = np.asarray([out_img.mean().item(), out_img.std().item(), out_img.abs().max().item()])
        for k in [f"block{i}" for i in range(case["depth"])] + ["final"]:
            v = taps[k]
            out[k + "_sample"] = v.numpy()[:, ::cases.TAP_TOKEN_STRIDE, ::cases.TAP_CHANNEL_STRIDE]
            out[k + "_moments"] = np.asarray([v.mean().item(), v.std().item(), v.abs().max().item()])
        # oracle restatement must agree with the real reference before fixtures are trusted
        o_img, o_te = orc.OracleDenoiser(st, depth=case["depth"])(img, t, x_t)
        err = ((o_te - out_te).abs().max() / out_te.abs().max()).item(), ((o_img - out_img).abs().max() / out_img.abs().max()).item()
        print(f"forward[{name}] oracle-vs-reference max err / max|ref|: te={err[0]:.3e} img={err[1]:.3e}")
        assert err[0] < 2e-5 and err[1] < 2e-5, err
        np.savez_compressed(os.path.join(OUT, f"forward_{name}.npz"), **out)


def golden_sampling(only=None):
    for name, case in cases.SAMPLING_CASES.items():
        if only and name not in only:
            continue
        m, st = build_ref(case)
        d = create_diffusion(case["respacing"])
        cond, noise = cases.sampling_inputs(case)
        torch.manual_seed(case["loop_seed"])   # the loop's randn_like stream (gaussian_diffusion.py:424)
        outs = list(d.p_sample_loop_progressive(m.forward, cond, noise.shape, noise, clip_denoised=False,
                                                model_kwargs=None, device="cpu", progress=False))
        keep = cases.kept_steps(len(outs))
        out = {"final": outs[-1]["sample"].numpy()}
        for n in keep:
            out[f"step{n}_sample"] = outs[n]["sample"].numpy()
            out[f"step{n}_x0"] = outs[n]["pred_xstart"].numpy()
        # metamorphic pin: the loop result is ONE forward at t=0 on the initial noise
        _, direct = m(cond, torch.zeros(noise.shape[0], dtype=torch.long), noise)
        out["quirk_maxdiff"] = np.asarray((direct - outs[-1]["sample"]).abs().max().item())
        # assignment on the final latents through the reference's own snippet (inference.py:294-306)
        G, tok = case["grid"], case["size"] // (16 * case["grid"])
        canon = torch.tensor(ref_models.get_2d_sincos_pos_embed(8, G)).unsqueeze(0).float()
        orders, preds, dists = [], [], []
        from einops import rearrange
        for b in range(noise.shape[0]):
            s = rearrange(outs[-1]["sample"][b], "(p1 h1 p2 w1) d -> (p1 p2) (h1 w1) d", p1=G, p2=G, h1=tok, w1=tok).mean(1)
            dist = pairwise_distances(s.cpu().numpy(), canon[0].cpu().numpy(), metric="manhattan")
            order = ref_inference.find_permutation(dist)
            orders.append(order); preds.append(np.asarray(order).argsort()); dists.append(dist)
        out["order"] = np.asarray(orders); out["pred"] = np.asarray(preds); out["dist"] = np.asarray(dists)
        # oracle restatement check
        sched = orc.Schedule(case["respacing"])
        torch.manual_seed(case["loop_seed"])
        step_noise = [torch.randn_like(noise) for _ in range(sched.num_timesteps)]
        o = list(sched.p_sample_loop_progressive(orc.OracleDenoiser(st, depth=case["depth"]), cond, noise, step_noise))
        e1 = max((o[n]["sample"] - outs[n]["sample"]).abs().max().item() for n in keep)
        print(f"sampling[{name}] oracle-vs-reference max abs err {e1:.3e}; quirk diff {out['quirk_maxdiff']}")
        assert e1 < 5e-5
        np.savez_compressed(os.path.join(OUT, f"sampling_{name}.npz"), **out)


def golden_ddim():
    """The reference's DDIM loop (gaussian_diffusion.py:531-578,636-698) run as committed, with the one argument its
    p_mean_variance call forgets (`condition`, :546) supplied by wrapping that bound method on the instance."""
    for name, case in cases.DDIM_CASES.items():
        m, st = build_ref(case)
        d = create_diffusion(case["respacing"])
        cond, noise = cases.sampling_inputs(case)
        inner = d.p_mean_variance
        d.p_mean_variance = lambda model, x, t, **kw: inner(model, cond, x, t, **kw)
        torch.manual_seed(case["loop_seed"])   # ddim_sample's randn_like stream (gaussian_diffusion.py:569)
        outs = list(d.ddim_sample_loop_progressive(m.forward, noise.shape, noise=noise, clip_denoised=False,
                                                   model_kwargs=None, device="cpu", progress=False, eta=case["eta"]))
        keep = cases.kept_steps(len(outs))
        out = {"final": outs[-1]["sample"].numpy()}
        for n in keep:
            out[f"step{n}_sample"] = outs[n]["sample"].numpy()
            out[f"step{n}_x0"] = outs[n]["pred_xstart"].numpy()
        sched = orc.Schedule(case["respacing"])
        torch.manual_seed(case["loop_seed"])
        step_noise = [torch.randn_like(noise) for _ in range(sched.num_timesteps)]
        model = orc.OracleDenoiser(st, depth=case["depth"])
        x, e1 = noise, 0.0
        for k, i in enumerate(range(sched.num_timesteps - 1, -1, -1)):
            o = sched.ddim_step(model, cond, x, torch.full((noise.shape[0],), i, dtype=torch.long), step_noise[k], case["eta"])
            x = o["sample"]
            if k in keep:
                e1 = max(e1, (o["sample"] - outs[k]["sample"]).abs().max().item(), (o["pred_xstart"] - outs[k]["pred_xstart"]).abs().max().item())
        print(f"ddim[{name}] oracle-vs-reference max abs err {e1:.3e} over {len(outs)} steps")
        assert e1 < 5e-5
        np.savez_compressed(os.path.join(OUT, f"ddim_{name}.npz"), **out)


def golden_training():
    for name, case in cases.TRAINING_CASES.items():
        m, st = build_ref(case)
        for p in m.parameters():
            p.requires_grad_(True)
        m.pos_embed.requires_grad_(False)
        d = create_diffusion("")
        x, t, piece = cases.training_inputs(case)
        G, bs = case["grid"], case["size"] // case["grid"]
        draws = cases.training_draws(case)     # replays the reference's RNG order on the same seeds
        cases.seed_training_rngs(case)
        with torch.enable_grad():
            terms = d.training_losses(m, x, t, piece, None, block_size=bs, patch_size=16,
                                      add_mask=case["add_mask"], grid_size=G)
            terms["loss"].mean().backward()
        out = {"mse": terms["mse"].detach().numpy(), "loss": terms["loss"].detach().numpy()}
        for k in cases.GRAD_KEYS:
            g = dict(m.named_parameters())[k].grad
            out["grad_norm/" + k] = np.asarray(g.norm().item())
            out["grad_head/" + k] = g.reshape(-1)[:64].numpy().copy()
        o = orc.training_losses(orc.Schedule(""), orc.OracleDenoiser(st, depth=case["depth"]), x, t, piece,
                                draws["perm"], draws["noise_x"], draws["noise_te"], block_size=bs, grid=G,
                                masks=draws["masks"])
        err = (o["mse"] - terms["mse"].detach()).abs().max().item()
        print(f"training[{name}] oracle-vs-reference mse err {err:.3e}  mse={out['mse']}")
        assert err < 1e-5
        np.savez_compressed(os.path.join(OUT, f"training_{name}.npz"), **out)
    # fresh-init known answer: mse == 0.5 exactly
    m = ref_models.DiT_models["JPDVT"](input_size=96, depth=1) if False else ref_models.DiT(
        input_size=96, depth=1, hidden_size=768, patch_size=16, num_heads=12)
    d = create_diffusion("")
    piece = torch.tensor(ref_models.get_2d_sincos_pos_embed(8, 3)).unsqueeze(0).float()
    torch.manual_seed(0); np.random.seed(0)
    terms = d.training_losses(m, torch.randn(2, 3, 96, 96), torch.tensor([3, 700]), piece, None,
                              block_size=32, patch_size=16, add_mask=False, grid_size=3)
    np.savez_compressed(os.path.join(OUT, "training_fresh.npz"), mse=terms["mse"].numpy())
    print("fresh-init mse", terms["mse"].numpy())


def golden_assignment():
    rng = np.random.default_rng(7)
    out = {}
    mats = []
    for n in (9, 16):
        for k in range(24):
            a = rng.standard_normal((n, 8)).astype(np.float32)
            b = ref_models.get_2d_sincos_pos_embed(8, int(n ** 0.5)).astype(np.float32)
            mats.append(pairwise_distances(a, b, metric="manhattan"))
    # adversarial score matrices: exact ties, values above the sentinel, duplicated rows, inf, nan
    for n in (9, 16):
        m = np.round(rng.standard_normal((n, n)) * 2) / 2          # many exact ties
        mats.append(np.abs(m))
        m = rng.random((n, n)) * 3e9                                 # entries above the 1e9 sentinel
        mats.append(m)
        m = rng.random((n, n)); m[3] = m[5]                          # duplicated rows
        mats.append(m)
        m = rng.random((n, n)); m[:, 2] = np.inf                     # a column of inf
        mats.append(m)
        m = rng.random((n, n)); m[4, 1] = np.nan; m[7, 6] = np.nan   # NaNs
        mats.append(m)
        mats.append(np.zeros((n, n)))                                # all equal
    for sentinel, tag in ((1e9, "1e9"), (2024.0, "2024")):
        for idx, mat in enumerate(mats):
            tmp = np.copy(mat)
            # the reference hard-codes its sentinel (1e9 in inference.py:124, 2024 in sample.py:103 /
            # train_JPDVT.py:556); the 2024 variant is the same loop with the constant swapped.
            if sentinel == 1e9:
                order = ref_inference.find_permutation(tmp)
            else:
                order, t2 = [], np.copy(mat)
                for _ in range(t2.shape[1]):
                    o = t2[:, 0].argmin(); order.append(o); t2 = t2[:, 1:]; t2[o, :] = 2024
            out[f"order_{tag}_{idx}"] = np.asarray(order, dtype=np.int64)
            out[f"pred_{tag}_{idx}"] = np.asarray(order).argsort()
            mine = orc.greedy_order(mat, sentinel)
            assert list(mine) == [int(v) for v in order], (tag, idx, mine, order)
    for idx, mat in enumerate(mats):
        out[f"scores_{idx}"] = mat
    out["n"] = np.asarray(len(mats))
    # l1 score restatement vs sklearn (bit-exact)
    a = rng.standard_normal((9, 8)).astype(np.float32)
    b = ref_models.get_2d_sincos_pos_embed(8, 3).astype(np.float32)
    assert np.array_equal(orc.l1_scores(a, b), pairwise_distances(a, b, metric="manhattan"))
    # perfect-latent round trip
    np.savez_compressed(os.path.join(OUT, "assignment.npz"), **out)
    print("assignment.npz", len(mats), "matrices x 2 sentinels")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["static", "assignment", "forward", "training", "sampling", "ddim"]
    for w in which:
        if ":" in w:                                   # e.g. sampling:c4_256g4_s250,c5_288_miss_s250 - regenerate named cases only
            fn, names = w.split(":", 1)
            globals()["golden_" + fn](set(names.split(",")))
        else:
            globals()["golden_" + w]()
