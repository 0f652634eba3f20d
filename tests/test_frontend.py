"""Batched puzzle front-end (jpdvt_mt_ntnu_b200/frontend.py): host logic on CPU, device kernels + solver on a B200.

The reference snippets being mirrored: scramble inference_ddp.py:382-395, scoring :431-447, reconstruction :449-455,
progress CSV :217-259, rank partition :325.  Scramble / reconstruct / score are pure data movement and integer compares:
bit-exact against the oracle (einops restatement) and numpy.
"""
import csv
import os

import numpy as np
import pytest
import torch

from oracle import cases
from oracle import jpdvt_oracle as orc


# ------------------------------------------------------------------------------------------------------------- CPU
def test_progress_csv_roundtrip_and_resume(tmp_path):
    from jpdvt_mt_ntnu_b200 import frontend as fe
    path = str(tmp_path / "logs" / "progress.csv")
    assert fe.load_progress_csv(path) == (set(), 0, 0, 0)                      # missing file: empty progress
    fe.append_progress_csv(path, [("a.jpg", 1, 9, 1.234), ("b.png", 0, 4, 0.5)])
    fe.append_progress_csv(path, [("c.JPEG", True, np.int32(9), 2.0)])
    fe.append_progress_csv(path, [])
    rows = list(csv.DictReader(open(path)))
    assert list(rows[0].keys()) == ["filename", "puzzle_correct", "patch_matches", "time_s"]   # the reference's columns
    assert [r["time_s"] for r in rows] == ["1.23", "0.50", "2.00"]             # f"{elapsed:.2f}"
    assert open(path).read().count("filename") == 1                            # header written once
    done, puzzles, pieces, count = fe.load_progress_csv(path)
    assert done == {"a.jpg", "b.png", "c.JPEG"} and (puzzles, pieces, count) == (2, 22, 3)


def test_list_images_and_center_crop(tmp_path):
    from PIL import Image
    from jpdvt_mt_ntnu_b200 import frontend as fe
    (tmp_path / "sub").mkdir()
    rng = np.random.RandomState(0)
    for name, (w, h) in {"b.jpg": (500, 300), "sub/a.png": (97, 230), "skip.txt": (8, 8), "c.JPEG": (192, 192)}.items():
        if name.endswith(".txt"):
            (tmp_path / name).write_text("x")
        else:
            Image.fromarray(rng.randint(0, 255, (h, w, 3), dtype=np.uint8)).save(tmp_path / name)
    found = fe.list_images(str(tmp_path))
    assert [os.path.relpath(p, tmp_path) for p in found] == ["b.jpg", "c.JPEG", "sub/a.png"]
    for p in found:
        t = fe.load_image(p, 96)
        assert tuple(t.shape) == (3, 96, 96) and t.dtype == torch.float32 and -1.0 <= float(t.min()) and float(t.max()) <= 1.0
    # an image that is already the target size passes through untouched: (u8 / 255 - 0.5) / 0.5
    arr = np.asarray(Image.open(tmp_path / "c.JPEG").convert("RGB"))
    want = (torch.from_numpy(arr.copy()).permute(2, 0, 1).float() / 255 - 0.5) / 0.5
    assert torch.equal(fe.load_image(str(tmp_path / "c.JPEG"), 192), want)


def test_solver_rejects_ill_formed_grid():
    from jpdvt_mt_ntnu_b200 import frontend as fe
    from jpdvt_mt_ntnu_b200.models import DiT
    m = DiT(input_size=288, depth=1, hidden_size=768, patch_size=16, num_heads=12)
    with pytest.raises(ValueError):                    # 4x4 @288: 72-px pieces are 4.5 tokens
        fe.PuzzleSolver(m, 4)


class _FakeSolver:
    """Duck-typed PuzzleSolver for the micro-batcher's host logic: 'predicts' the truth for even-sum images."""
    grid = 3

    def __init__(self):
        self.calls = []

    def draw_indices(self, batch):
        return np.stack([np.roll(np.arange(9), b + 1) for b in range(batch)]).astype(np.int32)

    def solve(self, images, indices=None, want_images=False, prescrambled=False, graph=False, **kw):
        from jpdvt_mt_ntnu_b200.frontend import SolveResult
        B = images.shape[0]
        self.calls.append((B, prescrambled))
        self.graph_calls = getattr(self, "graph_calls", 0) + int(graph)
        idx = torch.as_tensor(indices if indices is not None else self.draw_indices(B), dtype=torch.int32)
        pred = idx.clone()
        wrong = images.reshape(B, -1)[:, 0] < 0                 # requests flagged by a negative first pixel are mis-solved
        pred[wrong] = pred[wrong].roll(1, dims=1)
        eq = pred == idx
        return SolveResult(indices=idx, pred=pred, order=pred.argsort(1).int(), puzzle_correct=eq.all(1).int(),
                           patch_matches=eq.sum(1).int(), latents=torch.zeros(B, 1, 8), scrambled=images, reconstructed=images + 1)


def test_microbatcher_groups_requests_and_routes_results():
    import threading
    from jpdvt_mt_ntnu_b200 import frontend as fe
    fake = _FakeSolver()
    with fe.MicroBatcher(fake, max_batch=8, max_wait_ms=200.0) as mb:
        futs = {}
        def client(i):
            img = torch.full((3, 4, 4), float(i) if i % 3 else -1.0 - i)
            futs[i] = mb.submit(img, indices=np.roll(np.arange(9), i))
        threads = [threading.Thread(target=client, args=(i,)) for i in range(12)]
        [t.start() for t in threads]; [t.join() for t in threads]
        res = {i: f.result(timeout=30) for i, f in futs.items()}
    assert sum(b for b, _ in fake.calls) == 12 and max(b for b, _ in fake.calls) <= 8 and len(fake.calls) < 12   # batched
    for i, r in res.items():                                     # every caller gets ITS puzzle back
        assert r["indices"] == np.roll(np.arange(9), i).tolist() and r["total_patches"] == 9
        assert r["puzzle_correct"] == (1 if i % 3 else 0) and r["patch_accuracy"] == r["patch_matches"] / 9
        assert float(r["solution_image"][0, 0, 0]) == (float(i) if i % 3 else -1.0 - i) + 1
    # prescrambled requests need their indices, run as their own solve call, and errors reach the caller
    with fe.MicroBatcher(fake, max_batch=4, max_wait_ms=20.0) as mb:
        with pytest.raises(ValueError):
            mb.submit(torch.zeros(3, 4, 4), prescrambled=True)
        a = mb.submit(torch.ones(3, 4, 4), indices=np.arange(9), prescrambled=True)
        b = mb.submit(torch.ones(3, 4, 4))
        assert a.result(timeout=30)["puzzle_correct"] == 1 and b.result(timeout=30)["total_patches"] == 9
        fake.solve = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("boom"))
        with pytest.raises(RuntimeError):
            mb.submit(torch.ones(3, 4, 4)).result(timeout=30)
    with pytest.raises(RuntimeError):
        mb.submit(torch.ones(3, 4, 4))                           # closed
    # graph mode: every batch padded to max_batch, padding rows never reach a caller
    fake2 = _FakeSolver()
    with fe.MicroBatcher(fake2, max_batch=4, max_wait_ms=20.0, graph=True) as mb:
        outs = [mb.submit(torch.full((3, 4, 4), float(i + 1)), indices=np.roll(np.arange(9), i)).result(timeout=30) for i in range(3)]
    assert all(b == 4 for b, _ in fake2.calls) and fake2.graph_calls == len(fake2.calls)
    assert [o["indices"] for o in outs] == [np.roll(np.arange(9), i).tolist() for i in range(3)]


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("size,grid,batch", [(192, 3, 5), (256, 4, 3), (288, 3, 2), (192, 4, 2), (90, 3, 2), (48, 3, 1)])
def test_gather_pieces_matches_reference_scramble(cuda, size, grid, batch):
    from jpdvt_mt_ntnu_b200 import ops
    g = torch.Generator().manual_seed(size + grid)
    img = torch.rand(batch, 3, size, size, generator=g) * 2 - 1
    rs = np.random.RandomState(size)
    perms = np.stack([rs.permutation(grid * grid) for _ in range(batch)]).astype(np.int32)
    want = torch.cat([orc.scramble(img[b:b + 1], perms[b], grid) for b in range(batch)])
    got = ops.gather_pieces(img.cuda(), torch.from_numpy(perms).cuda(), grid)
    assert torch.equal(got.cpu(), want)                                         # data movement: bit-exact
    # masked-puzzle inference: selected slots zeroed, the others untouched
    keep = (rs.rand(batch, grid * grid) > 0.3).astype(np.uint8)
    got_m = ops.gather_pieces(img.cuda(), torch.from_numpy(perms).cuda(), grid, keep=torch.from_numpy(keep).cuda()).cpu()
    p = size // grid
    for b in range(batch):
        for s in range(grid * grid):
            ys, xs = (s // grid) * p, (s % grid) * p
            blk, ref = got_m[b, :, ys:ys + p, xs:xs + p], want[b, :, ys:ys + p, xs:xs + p]
            assert torch.equal(blk, ref if keep[b, s] else torch.zeros_like(ref))
    # reconstruction with a perfect prediction (pred == indices -> order = argsort(indices)) restores the image
    order = np.argsort(perms, axis=1).astype(np.int32)
    assert torch.equal(ops.gather_pieces(got, torch.from_numpy(order).cuda(), grid).cpu(), img)


@pytest.mark.gpu
def test_gather_pieces_rejects_bad_arguments(cuda):
    from jpdvt_mt_ntnu_b200 import ops
    from jpdvt_mt_ntnu_b200._lib import JpdvtError
    img = torch.zeros(2, 3, 100, 100, device="cuda")
    with pytest.raises(JpdvtError):                    # 100 px is not a multiple of a 3x3 grid
        ops.gather_pieces(img, torch.zeros(2, 9, dtype=torch.int32, device="cuda"), 3)
    with pytest.raises(JpdvtError):                    # wrong permutation shape
        ops.gather_pieces(img, torch.zeros(2, 9, dtype=torch.int32, device="cuda"), 5)
    assert ops.gather_pieces(img[:0], torch.zeros(0, 25, dtype=torch.int32, device="cuda"), 5).shape[0] == 0   # empty batch


@pytest.mark.gpu
@pytest.mark.parametrize("grid,in_piece,out_piece", [(3, 96, 64), (4, 64, 48), (3, 33, 20), (2, 10, 10), (3, 7, 4)])
def test_crop_pieces_matches_reference_erosion(cuda, grid, in_piece, out_piece):
    """train_JPDVT.py:345-349: rearrange -> torchvision CenterCrop -> rearrange, bit-exact (pure data movement)."""
    from einops import rearrange
    from torchvision import transforms
    from jpdvt_mt_ntnu_b200 import ops
    g = torch.Generator().manual_seed(grid * in_piece)
    x = torch.rand(3, 3, grid * in_piece, grid * in_piece, generator=g)
    patchs = rearrange(x, "b c (p1 h1) (p2 w1)-> b c (p1 p2) h1 w1", p1=grid, p2=grid, h1=in_piece, w1=in_piece)
    patchs = transforms.CenterCrop((out_piece, out_piece))(patchs)
    want = rearrange(patchs, "b c (p1 p2) h1 w1-> b c (p1 h1) (p2 w1)", p1=grid, p2=grid, h1=out_piece, w1=out_piece)
    assert torch.equal(ops.crop_pieces(x.cuda(), grid, out_piece).cpu(), want)


@pytest.mark.gpu
def test_score_placements(cuda):
    from jpdvt_mt_ntnu_b200 import ops
    rs = np.random.RandomState(3)
    for n in (9, 16, 25, 36):
        truth = np.stack([rs.permutation(n) for _ in range(37)]).astype(np.int32)
        pred = truth.copy()
        for b in range(0, 37, 2):                       # corrupt every other puzzle
            i, j = rs.choice(n, 2, replace=False)
            pred[b, [i, j]] = pred[b, [j, i]]
        totals = torch.zeros(3, dtype=torch.int64, device="cuda")
        correct, matches = ops.score_placements(torch.from_numpy(pred).cuda(), torch.from_numpy(truth).cuda(), totals)
        eq = pred == truth
        assert np.array_equal(matches.cpu().numpy(), eq.sum(1)) and np.array_equal(correct.cpu().numpy(), eq.all(1).astype(np.int32))
        assert totals.tolist() == [int(eq.all(1).sum()), int(eq.sum()), 37]
        ops.score_placements(torch.from_numpy(pred).cuda(), torch.from_numpy(truth).cuda(), totals)            # accumulates
        assert totals.tolist() == [2 * int(eq.all(1).sum()), 2 * int(eq.sum()), 74]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny48_s10", "d2_192_s250", "d2_256g4_s25"])
def test_solver_reproduces_reference_pipeline(cuda, golden, name):
    """PuzzleSolver.solve on the UNSCRAMBLED images reproduces the reference per-image pipeline end to end: the scramble is
    bit-identical to the golden case's condition, the placements equal the reference's `pred` (golden fixture)."""
    from jpdvt_mt_ntnu_b200 import frontend as fe
    from jpdvt_mt_ntnu_b200.models import DiT
    case = cases.SAMPLING_CASES[name]
    gold = golden("sampling_" + name)
    m = DiT(input_size=case["size"], depth=case["depth"], hidden_size=768, patch_size=16, num_heads=12)
    m.load_state_dict(cases.state_for(case))
    m.cuda()
    steps = int(case["respacing"])
    solver = fe.PuzzleSolver(m, case["grid"], sampling_steps=steps)
    # the case's own inputs: images (before scrambling), permutations, the shared noise row, the loop's randn stream
    gen = torch.Generator().manual_seed(case["seed"])
    img = torch.rand(case["batch"], 3, case["size"], case["size"], generator=gen) * 2 - 1
    perms = np.stack(cases.sampling_perms(case)).astype(np.int32)
    cond, noise = cases.sampling_inputs(case)
    solver.noise_row = noise[:1].cuda()
    torch.manual_seed(case["loop_seed"])
    step_noise = torch.stack([torch.randn_like(noise) for _ in range(steps)]).cuda()
    res = solver.solve(img.pin_memory(), indices=perms, step_noise=step_noise, want_images=True)
    assert torch.equal(res.scrambled.cpu(), cond)
    assert np.array_equal(res.pred.cpu().numpy(), gold["pred"]) and np.array_equal(res.order.cpu().numpy(), gold["order"])
    eq = gold["pred"] == perms
    assert np.array_equal(res.patch_matches.cpu().numpy(), eq.sum(1)) and np.array_equal(res.puzzle_correct.cpu().numpy(), eq.all(1))
    assert solver.running_totals() == (int(eq.all(1).sum()), int(eq.sum()), case["batch"])
    # reconstruction: cell j shows scrambled slot order[j] (inference_ddp.py:449-455)
    p, G = case["size"] // case["grid"], case["grid"]
    for b in range(case["batch"]):
        for j in range(G * G):
            s = int(gold["order"][b][j])
            a = res.reconstructed[b, :, (j // G) * p:(j // G + 1) * p, (j % G) * p:(j % G + 1) * p].cpu()
            assert torch.equal(a, cond[b, :, (s // G) * p:(s // G + 1) * p, (s % G) * p:(s % G + 1) * p])


@pytest.mark.gpu
def test_solve_files_csv_resume_and_masking(cuda, tmp_path):
    from PIL import Image
    from jpdvt_mt_ntnu_b200 import frontend as fe
    from jpdvt_mt_ntnu_b200.models import DiT
    rng = np.random.RandomState(1)
    for i in range(7):
        Image.fromarray(rng.randint(0, 255, (120 + 10 * i, 140, 3), dtype=np.uint8)).save(tmp_path / f"im{i}.png")
    (tmp_path / "broken.png").write_bytes(b"not an image")
    case = cases.SAMPLING_CASES["tiny48_s10"]
    m = DiT(input_size=96, depth=2, hidden_size=768, patch_size=16, num_heads=12)
    m.load_state_dict(orc.seeded_state(m.state_dict(), seed=3))
    m.cuda()
    solver = fe.PuzzleSolver(m, 3, sampling_steps=5, missing_per_puzzle=(1, 2))
    paths = fe.list_images(str(tmp_path))
    csv_path = str(tmp_path / "out" / "progress.csv")
    fe.append_progress_csv(csv_path, [("im3.png", 1, 9, 0.1)])                   # an earlier run already did im3
    seen = []
    stats = fe.solve_files(solver, paths, csv_path, batch_size=4, on_batch=lambda names, res: seen.extend(names))
    assert sorted(seen) == [f"im{i}.png" for i in (0, 1, 2, 4, 5, 6)]            # resume skipped im3; broken.png skipped
    done, puzzles, pieces, count = fe.load_progress_csv(csv_path)
    assert count == 7 and done == {f"im{i}.png" for i in range(7)}
    assert stats["puzzles"] == 7 and 0.0 <= stats["patch_accuracy"] <= 1.0       # resumed counters folded in
    assert fe.solve_files(solver, paths, csv_path, batch_size=4)["puzzles"] == 7  # nothing left to do: totals from the CSV


@pytest.mark.gpu
def test_microbatcher_and_prescrambled_on_device(cuda):
    from jpdvt_mt_ntnu_b200 import frontend as fe, ops
    from jpdvt_mt_ntnu_b200.models import DiT
    m = DiT(input_size=96, depth=2, hidden_size=768, patch_size=16, num_heads=12)
    m.load_state_dict(orc.seeded_state(m.state_dict(), seed=5))
    m.cuda()
    solver = fe.PuzzleSolver(m, 3, sampling_steps=4)
    g = torch.Generator().manual_seed(2)
    imgs = torch.rand(6, 3, 96, 96, generator=g) * 2 - 1
    perms = np.stack([np.random.RandomState(i).permutation(9) for i in range(6)]).astype(np.int32)
    # prescrambled: the images pass through untouched and are scored against the given truth
    pre = ops.gather_pieces(imgs.cuda(), torch.from_numpy(perms).cuda(), 3)
    torch.manual_seed(0)
    r1 = solver.solve(pre, indices=perms, prescrambled=True, want_images=True)
    torch.manual_seed(0)
    r2 = solver.solve(imgs, indices=perms, want_images=True)
    assert torch.equal(r1.scrambled, pre) and torch.equal(r2.scrambled, pre)
    assert torch.equal(r1.pred, r2.pred) and torch.equal(r1.patch_matches, r2.patch_matches)      # same puzzles, same loop noise
    with pytest.raises(ValueError):
        solver.solve(pre, prescrambled=True)
    with fe.MicroBatcher(solver, max_batch=4, max_wait_ms=50.0) as mb:
        futs = [mb.submit(imgs[i], indices=perms[i]) for i in range(6)]
        out = [f.result(timeout=120) for f in futs]
    for i, r in enumerate(out):
        assert r["indices"] == perms[i].tolist() and sorted(r["predicted_order"]) == list(range(9))
        assert torch.equal(r["scrambled_image"], pre[i]) and tuple(r["solution_image"].shape) == (3, 96, 96)
        assert r["patch_matches"] == int((np.asarray(r["predicted_order"]) == perms[i]).sum())
    assert 2 <= mb.batches_run <= 6
    # the whole loop replayed from a CUDA graph: bit-identical to the launched loop
    torch.manual_seed(0)
    r3 = solver.solve(imgs, indices=perms, graph=True)
    torch.manual_seed(0)
    r4 = solver.solve(imgs, indices=perms, graph=True)                                            # replay of the cached graph
    assert torch.equal(r3.latents, r2.latents) and torch.equal(r4.latents, r2.latents) and torch.equal(r3.pred, r2.pred)
