"""The CPU oracle restatement against the fixtures produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import cases
from oracle import jpdvt_oracle as orc

TOL = 2e-5   # fp32 CPU vs fp32 CPU, different op order (conv vs matmul, SDPA vs explicit softmax)


def _close(a, b, tol=TOL):
    scale = max(float(np.abs(b).max()), 1e-6)
    assert float(np.abs(a - b).max()) <= tol * scale, (float(np.abs(a - b).max()), scale)


def test_pos_embed_known_answers(golden):
    g = golden("static")
    assert np.array_equal(orc.sincos_2d(8, 3), g["pe_8_3"])
    assert np.array_equal(orc.sincos_2d(8, 4), g["pe_8_4"])
    assert np.array_equal(orc.sincos_2d(768, 12), g["pe_768_12"])
    # row 1 of get_2d_sincos_pos_embed(8, 3)
    np.testing.assert_allclose(orc.sincos_2d(8, 3)[1], [0.84147098, 0.00999983, 0.54030231, 0.99995, 0, 0, 1, 1], atol=1e-8)


@pytest.mark.parametrize("name,spec", [("full", ""), ("s250", "250"), ("s10", "10"), ("ddim50", "ddim50"), ("sec", "10,15,20")])
def test_schedule_tables_bit_exact(golden, name, spec):
    g = golden("static")
    s = orc.Schedule(spec)
    assert np.array_equal(np.asarray(s.timestep_map), g[f"{name}_map"])
    for key, arr in (("betas", s.betas), ("sqrt_ac", s.sqrt_ac), ("sqrt_1mac", s.sqrt_1mac), ("post_var", s.post_var),
                     ("post_logvar", s.post_logvar), ("coef1", s.coef1), ("coef2", s.coef2)):
        assert np.array_equal(arr, g[f"{name}_{key}"]), key


def test_schedule_known_answers():
    s = orc.Schedule("250")
    assert s.timestep_map[:3] == [0, 4, 8] and s.timestep_map[-3:] == [991, 995, 999]
    assert s.coef1[0] == 1.0 and s.coef2[0] == 0.0
    assert abs(s.betas[1] - 5.990655644756426e-04) < 1e-15 and abs(s.betas[-1] - 0.07751934499235003) < 1e-15
    full = orc.Schedule("")
    assert abs(full.alphas_cumprod[-1] - 4.035829765375676e-05) < 1e-18


def test_state_layout_matches_reference(golden):
    g = golden("static")
    st = orc.blank_state(192, 12)
    assert list(st.keys()) == [str(k) for k in g["state_keys"]]
    assert [v.numel() for v in st.values()] == g["state_numel"].tolist()
    assert sum(v.numel() for v in st.values()) == int(g["n_params_192"]) == 130857800


@pytest.mark.parametrize("name", ["tiny48", "d2_192", "d2_256", "d2_288", "hot192", "full192"])
def test_forward_matches_reference(golden, name):
    case = cases.FORWARD_CASES[name]
    g = golden("forward_" + name)
    img, t, x_t = cases.forward_inputs(case)
    taps = {}
    with torch.no_grad():
        o_img, o_te = orc.OracleDenoiser(cases.state_for(case), depth=case["depth"])(img, t, x_t, taps)
    _close(o_te.numpy(), g["te"])
    _close(taps["c"].numpy(), g["c"])
    _close(o_img.numpy()[:, :, ::7, ::5], g["img_sample"])
    for k in [f"block{i}" for i in range(case["depth"])] + ["final"]:
        _close(taps[k].numpy()[:, ::cases.TAP_TOKEN_STRIDE, ::cases.TAP_CHANNEL_STRIDE], g[k + "_sample"])


@pytest.mark.parametrize("name", ["tiny48_s10", "d2_256g4_s25"])
def test_sampling_loop_matches_reference(golden, name):
    case = cases.SAMPLING_CASES[name]
    g = golden("sampling_" + name)
    cond, noise = cases.sampling_inputs(case)
    sched = orc.Schedule(case["respacing"])
    torch.manual_seed(case["loop_seed"])
    step_noise = [torch.randn_like(noise) for _ in range(sched.num_timesteps)]
    with torch.no_grad():
        outs = list(sched.p_sample_loop_progressive(orc.OracleDenoiser(cases.state_for(case), depth=case["depth"]), cond, noise, step_noise))
    for n in cases.kept_steps(len(outs)):
        _close(outs[n]["sample"].numpy(), g[f"step{n}_sample"], 5e-5)
        _close(outs[n]["pred_xstart"].numpy(), g[f"step{n}_x0"], 5e-5)
    _close(outs[-1]["sample"].numpy(), g["final"], 5e-5)
    G, tok = case["grid"], case["size"] // (16 * case["grid"])
    for b in range(case["batch"]):
        order, pred, sc = orc.solve(torch.from_numpy(g["final"][b]), G, tok)
        assert np.array_equal(sc, g["dist"][b])            # fp64 L1 scores bit-exact vs sklearn on identical latents
        assert list(order) == g["order"][b].tolist() and list(pred) == g["pred"][b].tolist()


@pytest.mark.parametrize("name", ["d2_192_s250", "full192_s250", "c4_256g4_s250", "c5_288_miss_s250"])
def test_loop_quirk_one_forward_equals_loop(golden, name):
    """the reference loop result is ONE forward at t=0 on the initial noise (gaussian_diffusion.py:518-529)."""
    case = cases.SAMPLING_CASES[name]
    g = golden("sampling_" + name)
    assert float(g["quirk_maxdiff"]) == 0.0
    cond, noise = cases.sampling_inputs(case)
    with torch.no_grad():
        _, te = orc.OracleDenoiser(cases.state_for(case), depth=case["depth"])(cond, torch.zeros(case["batch"], dtype=torch.long), noise)
    _close(te.numpy(), g["final"], 5e-5)


@pytest.mark.parametrize("name", list(cases.TRAINING_CASES))
def test_training_losses_match_reference(golden, name):
    case = cases.TRAINING_CASES[name]
    g = golden("training_" + name)
    x, t, piece = cases.training_inputs(case)
    d = cases.training_draws(case)
    with torch.no_grad():
        o = orc.training_losses(orc.Schedule(""), orc.OracleDenoiser(cases.state_for(case), depth=case["depth"]), x, t, piece,
                                d["perm"], d["noise_x"], d["noise_te"], block_size=case["size"] // case["grid"],
                                grid=case["grid"], masks=d["masks"])
    np.testing.assert_allclose(o["mse"].numpy(), g["mse"], rtol=1e-5, atol=1e-6)


def test_fresh_init_mse_is_half(golden):
    assert golden("training_fresh")["mse"].tolist() == [0.5, 0.5]
    assert golden("static")["fresh_absmax"].tolist() == [0.0, 0.0]


def test_assignment_bit_exact_vs_reference(golden):
    g = golden("assignment")
    for tag, sentinel in (("1e9", 1e9), ("2024", 2024.0)):
        for i in range(int(g["n"])):
            order = orc.greedy_order(g[f"scores_{i}"], sentinel)
            assert order == g[f"order_{tag}_{i}"].tolist(), (tag, i)
            if sorted(order) == list(range(len(order))):
                assert orc.placements(order).tolist() == g[f"pred_{tag}_{i}"].tolist()


def test_perfect_latents_round_trip():
    """tokens set to the canonical embedding of their piece => pred == the scramble indices."""
    rs = np.random.RandomState(0)
    for G, tok in ((3, 4), (4, 4), (3, 6)):
        canon = torch.from_numpy(orc.sincos_2d(8, G)).float()
        perm = rs.permutation(G * G)
        lat = orc.expand_piece_embeddings(canon[perm][None], G, tok)[0]
        _, pred, _ = orc.solve(lat, G, tok)
        assert pred.tolist() == perm.tolist()


@pytest.mark.parametrize("name", sorted(cases.DDIM_CASES))
def test_ddim_loop_matches_reference(golden, name):
    """The reference's own DDIM code (gaussian_diffusion.py:531-578,636-698) with `condition` supplied at its
    p_mean_variance call (fixtures: oracle/make_golden.py golden_ddim) vs the oracle's ddim_step."""
    case = cases.DDIM_CASES[name]
    g = golden("ddim_" + name)
    cond, noise = cases.sampling_inputs(case)
    sched = orc.Schedule(case["respacing"])
    torch.manual_seed(case["loop_seed"])
    step_noise = [torch.randn_like(noise) for _ in range(sched.num_timesteps)]
    model = orc.OracleDenoiser(cases.state_for(case), depth=case["depth"])
    keep = cases.kept_steps(sched.num_timesteps)
    x = noise
    with torch.no_grad():
        for k, i in enumerate(range(sched.num_timesteps - 1, -1, -1)):
            o = sched.ddim_step(model, cond, x, torch.full((case["batch"],), i, dtype=torch.long), step_noise[k], case["eta"])
            x = o["sample"]
            if k in keep:
                _close(o["sample"].numpy(), g[f"step{k}_sample"], 5e-5)
                _close(o["pred_xstart"].numpy(), g[f"step{k}_x0"], 5e-5)
    _close(x.numpy(), g["final"], 5e-5)


def test_philox_restatement_known_answers():
    """oracle.philox4x32_10 against the published Random123 known-answer vectors for philox4x32-10 (kat_vectors: all-zero
    counter/key, all-ones counter/key, and the pi-digits vector), so the GPU test that compares the in-kernel step-noise
    stream with it is anchored outside this repo."""
    import numpy as np
    from oracle import jpdvt_oracle as orc
    z = orc.philox4x32_10(np.zeros((1, 4), dtype=np.uint32), (0, 0))[0]
    assert [int(v) for v in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    o = orc.philox4x32_10(np.full((1, 4), 0xffffffff, dtype=np.uint64), (0xffffffff, 0xffffffff))[0]
    assert [int(v) for v in o] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    p = orc.philox4x32_10(np.array([[0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], dtype=np.uint64), (0xa4093822, 0x299f31d0))[0]
    assert [int(v) for v in p] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    n = orc.philox_normals(1 << 16, seed=12345, call=0, step=3)
    assert abs(n.mean()) < 2e-2 and abs(n.std() - 1) < 2e-2
