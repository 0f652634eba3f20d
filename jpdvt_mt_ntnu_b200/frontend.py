"""Batched front-end for puzzle solving: scramble / mask / sample / assign / score / reconstruct for a whole batch of
puzzles on one GPU, one process per GPU.

It replaces the per-image Python loops of the reference inference scripts (image_model/inference_ddp.py:338-470,
inferencetexmet.py:296-405, which scrambles image by image on the host and fans a batch out with nn.DataParallel threads)
with device kernels on either side of the sampling loop, and keeps the files those scripts exchange:

  * the progress CSV (`filename,puzzle_correct,patch_matches,time_s`, inference_ddp.py:217-259) with its resume
    semantics - files already listed are skipped, their counters are folded into the totals;
  * the rank partition `image_paths[rank::world_size]` (inference_ddp.py:325) and the closing SUM / MAX all-reduce of
    (puzzles correct, pieces correct, count) / wall time (inference_ddp.py:485-495).

All arithmetic runs in libjpdvt_sm100.so (ops.gather_pieces, the sampling loop, ops.assign_greedy_l1,
ops.score_placements); this module only owns buffers, RNG streams and files.
"""
from __future__ import annotations

import csv
import os
import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import assignment, ops, parallel
from .diffusion import create_diffusion

CSV_FIELDS = ["filename", "puzzle_correct", "patch_matches", "time_s"]
ALLOWED_EXTENSIONS = (".jpg", ".jpeg", ".png", ".JPEG")          # inference_ddp.py:69


# ----------------------------------------------------------------------------------------------- progress CSV / resume
def load_progress_csv(csv_path: str) -> Tuple[set, int, int, int]:
    """inference_ddp.py:217-242: (processed filenames, puzzles correct, pieces correct, rows)."""
    done, puzzles, pieces, count = set(), 0, 0, 0
    if not os.path.exists(csv_path):
        return done, puzzles, pieces, count
    with open(csv_path, "r") as f:
        for row in csv.DictReader(f):
            done.add(row["filename"])
            puzzles += int(row["puzzle_correct"])
            pieces += int(row["patch_matches"])
            count += 1
    return done, puzzles, pieces, count


def append_progress_csv(csv_path: str, rows: Iterable[Tuple[str, int, int, float]]) -> None:
    """inference_ddp.py:244-259, one or many rows per call (one open/append per batch instead of per image)."""
    rows = list(rows)
    if not rows:
        return
    exists = os.path.exists(csv_path)
    os.makedirs(os.path.dirname(os.path.abspath(csv_path)), exist_ok=True)
    with open(csv_path, "a", newline="") as f:
        w = csv.DictWriter(f, fieldnames=CSV_FIELDS)
        if not exists:
            w.writeheader()
        for name, ok, matches, seconds in rows:
            w.writerow({"filename": name, "puzzle_correct": int(ok), "patch_matches": int(matches), "time_s": f"{seconds:.2f}"})


# ----------------------------------------------------------------------------------------------- image loading (host)
def center_crop_square(pil_image, image_size: int):
    """The ADM centre crop the reference scripts use (inference_ddp.py:173-189): halve with a box filter while the short
    side is >= 2x the target, bicubic-resize the short side to the target, crop the centre."""
    from PIL import Image
    while min(pil_image.size) >= 2 * image_size:
        pil_image = pil_image.resize(tuple(v // 2 for v in pil_image.size), resample=Image.BOX)
    scale = image_size / min(pil_image.size)
    pil_image = pil_image.resize(tuple(round(v * scale) for v in pil_image.size), resample=Image.BICUBIC)
    arr = np.asarray(pil_image)
    y0, x0 = (arr.shape[0] - image_size) // 2, (arr.shape[1] - image_size) // 2
    return arr[y0:y0 + image_size, x0:x0 + image_size]


def load_image(path: str, image_size: int) -> torch.Tensor:
    """RGB -> centre crop -> [3,S,S] fp32 in [-1,1]  (ToTensor + Normalize(0.5, 0.5), inference_ddp.py:280-284)."""
    from PIL import Image
    arr = center_crop_square(Image.open(path).convert("RGB"), image_size)
    return torch.from_numpy(np.array(arr, copy=True)).permute(2, 0, 1).float().div_(255.0).sub_(0.5).div_(0.5)


def list_images(data_dir: str, extensions: Sequence[str] = ALLOWED_EXTENSIONS) -> List[str]:
    """Recursive, sorted listing (inference_ddp.py:314-319)."""
    found = []
    for root, _, files in os.walk(data_dir):
        found.extend(os.path.join(root, f) for f in files if f.endswith(tuple(extensions)))
    return sorted(found)


# ----------------------------------------------------------------------------------------------- the batched solver
@dataclass
class SolveResult:
    indices: torch.Tensor          # int32 [B, G*G]  ground-truth scramble (slot i holds original piece indices[i])
    pred: torch.Tensor             # int32 [B, G*G]  predicted cell of every slot (np.asarray(order).argsort())
    order: torch.Tensor            # int32 [B, G*G]  find_permutation's sort_list
    puzzle_correct: torch.Tensor   # int32 [B]
    patch_matches: torch.Tensor    # int32 [B]
    latents: torch.Tensor          # fp32 [B, T, 8]   p_sample_loop result
    scrambled: Optional[torch.Tensor] = None       # fp32 [B,3,S,S]
    reconstructed: Optional[torch.Tensor] = None   # fp32 [B,3,S,S]


class PuzzleSolver:
    """One GPU's share of the puzzles: `solve(images)` runs the whole inference_ddp.py per-image body for a batch.

    model: a jpdvt_mt_ntnu_b200.models.DiT on the device; grid_size G; `sampling_steps` as create_diffusion(str(N)).
    RNG streams follow the reference: one randn(1,T,8) initial-noise row drawn at construction after
    torch.manual_seed(seed + rank) and shared by every puzzle (inference_ddp.py:278,311-313; repeated over the batch as
    inferencetexmet.py:313); scramble permutations from a numpy RandomState(seed + rank) unless passed in.
    """

    def __init__(self, model, grid_size: int, sampling_steps: int = 250, seed: int = 0, rank: int = 0, sentinel: float = 1e9,
                 missing_per_puzzle: Tuple[int, int] = (0, 0)):
        self.model, self.grid, self.sentinel = model, int(grid_size), float(sentinel)
        self.size = int(model.input_size) if hasattr(model, "input_size") else int(model.x_embedder.img_size[0])
        if self.size % (16 * self.grid) != 0:
            raise ValueError(f"{self.size}px / {self.grid}x{self.grid}: pieces must be whole 16-px tokens (inference.py:295)")
        self.tokens = (self.size // 16) ** 2
        self.device = next(model.parameters()).device
        self.diffusion = create_diffusion(str(sampling_steps))
        self.np_rng = np.random.RandomState(seed + rank)
        self.missing_per_puzzle = missing_per_puzzle
        gen = torch.Generator(device="cpu").manual_seed(seed + rank)
        self.noise_row = torch.randn(1, self.tokens, 8, generator=gen).to(self.device)
        self.totals = torch.zeros(3, dtype=torch.int64, device=self.device)      # puzzles correct, pieces correct, puzzles

    def draw_indices(self, batch: int) -> np.ndarray:
        return np.stack([self.np_rng.permutation(self.grid * self.grid) for _ in range(batch)]).astype(np.int32)

    def draw_missing(self, batch: int) -> Optional[np.ndarray]:
        """keep mask [B, G*G] uint8 with r in [lo, hi] random slots zeroed per puzzle (masked-puzzle inference, C5)."""
        lo, hi = self.missing_per_puzzle
        if hi <= 0:
            return None
        keep = np.ones((batch, self.grid * self.grid), dtype=np.uint8)
        for b in range(batch):
            r = int(self.np_rng.randint(lo, hi + 1))
            keep[b, self.np_rng.choice(self.grid * self.grid, size=r, replace=False)] = 0
        return keep

    @torch.no_grad()
    def solve(self, images: torch.Tensor, indices=None, keep=None, step_noise: Optional[torch.Tensor] = None,
              want_images: bool = False, prescrambled: bool = False, graph: bool = False,
              count_rows: Optional[int] = None) -> SolveResult:
        """images [B,3,S,S] fp32 in [-1,1] (host - pinned or not - or device).  indices [B,G*G] (default: drawn), keep
        [B,G*G] 0/1 (default: drawn from `missing_per_puzzle`, None = nothing missing).  prescrambled=True: the images
        ARE the puzzles (api/app.py:350-451 receives them that way); `indices` is then only the ground truth to score
        against (required).  graph=True replays the sampling loop from a CUDA graph (small, repeated batch shapes: -13 % at
        batch 1, -16 % at batch 16).  count_rows: only the first `count_rows` puzzles enter the running totals (the rest
        are shape padding of a graph-replayed batch)."""
        from . import _lib
        with _lib.on_device(self.device):
            return self._solve(images, indices, keep, step_noise, want_images, prescrambled, graph, count_rows)

    def _solve(self, images, indices, keep, step_noise, want_images, prescrambled, graph, count_rows):
        B = images.shape[0]
        n = self.grid * self.grid
        if tuple(images.shape[1:]) != (3, self.size, self.size):
            raise ValueError(f"expected images [B,3,{self.size},{self.size}], got {tuple(images.shape)}")
        x = images.to(self.device, dtype=torch.float32, non_blocking=True).contiguous()
        if indices is None:
            if prescrambled:
                raise ValueError("prescrambled puzzles need their scramble indices to be scored")
            indices = self.draw_indices(B)
        if keep is None and not prescrambled:
            keep = self.draw_missing(B)
        idx = torch.as_tensor(np.asarray(indices), dtype=torch.int32).reshape(B, n).to(self.device)
        keep_t = None if keep is None else torch.as_tensor(np.asarray(keep), dtype=torch.uint8).reshape(B, n).to(self.device)
        scrambled = x if prescrambled else ops.gather_pieces(x, idx, self.grid, keep=keep_t)
        noise = self.noise_row.expand(B, -1, -1).contiguous()
        latents = self.diffusion.p_sample_loop(self.model.forward, scrambled, noise.shape, noise, clip_denoised=False,
                                               model_kwargs=None, progress=False, device=self.device, step_noise=step_noise,
                                               graph=graph)
        order, pred = assignment.solve_puzzles(latents, self.grid, self.sentinel)
        if count_rows is None or count_rows >= B:
            correct, matches = ops.score_placements(pred, idx, totals=self.totals)
        else:
            correct, matches = ops.score_placements(pred, idx)
            if count_rows > 0:
                ops.score_placements(pred[:count_rows].contiguous(), idx[:count_rows].contiguous(), totals=self.totals)
        res = SolveResult(indices=idx, pred=pred, order=order, puzzle_correct=correct, patch_matches=matches, latents=latents)
        if want_images:
            res.scrambled = scrambled
            res.reconstructed = ops.gather_pieces(scrambled, order, self.grid)     # cell j shows slot order[j]
        return res

    def running_totals(self) -> Tuple[int, int, int]:
        t = self.totals.tolist()
        return int(t[0]), int(t[1]), int(t[2])


# ----------------------------------------------------------------------------------------------- directory driver
def solve_files(solver: PuzzleSolver, paths: Sequence[str], csv_path: str, batch_size: int = 256, rank: int = 0,
                world_size: int = 1, loader: Optional[Callable[[str, int], torch.Tensor]] = None,
                on_batch: Optional[Callable[[List[str], SolveResult], None]] = None) -> Dict[str, float]:
    """The inference_ddp.py main loop, batched: this rank's `paths[rank::world_size]`, minus files already in the progress
    CSV, in batches of `batch_size`; one CSV row per image; returns the global statistics after the closing all-reduce.
    Files that fail to load are skipped (the reference's per-image try/except, inference_ddp.py:340,466-469)."""
    loader = loader or load_image
    mine = parallel.strided_shard(list(paths), rank, world_size)
    done, puzzles0, pieces0, count0 = load_progress_csv(csv_path)
    todo = [p for p in mine if os.path.basename(p) not in done]
    # counters resume "proportionally" as the reference does (inference_ddp.py:332-335)
    puzzles, pieces, count = puzzles0 // world_size, pieces0 // world_size, count0 // world_size
    n = solver.grid * solver.grid
    t_start = time.time()
    staging = torch.empty(batch_size, 3, solver.size, solver.size).pin_memory()      # decoded batch, H2D straight from here
    for b0 in range(0, len(todo), batch_size):
        names, k = [], 0
        t0 = time.time()
        for p in todo[b0:b0 + batch_size]:
            try:
                img = loader(p, solver.size)
            except Exception:
                continue
            staging[k].copy_(img)
            names.append(os.path.basename(p)); k += 1
        if k == 0:
            continue
        res = solver.solve(staging[:k])
        ok, matches = res.puzzle_correct.tolist(), res.patch_matches.tolist()
        per_image = (time.time() - t0) / k
        append_progress_csv(csv_path, [(nm, o, m, per_image) for nm, o, m in zip(names, ok, matches)])
        puzzles += sum(ok); pieces += sum(matches); count += k
        if on_batch is not None:
            on_batch(names, res)
    (gp, gm, gc), wall = parallel.reduce_stats(puzzles, pieces, count, time.time() - t_start, solver.device)
    return {"puzzles": gc, "puzzle_accuracy": gp / gc if gc else 0.0, "patch_accuracy": gm / (gc * n) if gc else 0.0,
            "wall_s": wall, "puzzles_per_s": gc / wall if wall > 0 else 0.0}


# ----------------------------------------------------------------------------------------------- request micro-batching
class MicroBatcher:
    """Groups single-puzzle requests into batches for one PuzzleSolver: the serving path of
    api/app.py:250-451 solves one upload per call; here concurrent callers `submit()` an image and get a Future, a worker
    thread collects up to `max_batch` requests - or whatever arrived within `max_wait_ms` of the first - and runs them as
    ONE `solver.solve` call.  Each Future resolves to the per-request fields the API returns (`metrics` / `details`,
    app.py:335-345) plus the reconstructed image tensor.
    """

    def __init__(self, solver, max_batch: int = 64, max_wait_ms: float = 5.0, graph: bool = False):
        """graph=True: every batch is padded to `max_batch` (repeating its last request) and the sampling loop is replayed from
        one CUDA graph - a fixed shape, no host launches in the 250-step loop."""
        self.solver, self.max_batch, self.max_wait = solver, int(max_batch), float(max_wait_ms) / 1000.0
        self.graph = bool(graph)
        self._q: "queue.Queue" = queue.Queue()
        self._stop = threading.Event()
        self.batches_run = 0
        self._thread = threading.Thread(target=self._loop, name="jpdvt-microbatch", daemon=True)
        self._thread.start()

    def submit(self, image: torch.Tensor, indices=None, prescrambled: bool = False) -> Future:
        """image [3,S,S] fp32 in [-1,1]; indices: optional scramble ground truth (required when prescrambled)."""
        if self._stop.is_set():
            raise RuntimeError("MicroBatcher is closed")
        if prescrambled and indices is None:
            raise ValueError("prescrambled puzzles need their scramble indices")
        fut: Future = Future()
        self._q.put((image, None if indices is None else np.asarray(indices, dtype=np.int32), bool(prescrambled), fut))
        return fut

    def close(self) -> None:
        self._stop.set()
        self._q.put(None)
        self._thread.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _collect(self):
        first = self._q.get()
        if first is None:
            return None
        items, deadline = [first], time.monotonic() + self.max_wait
        while len(items) < self.max_batch:
            left = deadline - time.monotonic()
            if left <= 0:
                break
            try:
                nxt = self._q.get(timeout=left)
            except queue.Empty:
                break
            if nxt is None:
                self._q.put(None)           # leave the stop marker for the loop
                break
            items.append(nxt)
        return items

    def _run(self, items) -> None:
        n = self.solver.grid * self.solver.grid
        # one solve per kind: scrambling is a per-batch switch in PuzzleSolver.solve
        for kind in (False, True):
            group = [it for it in items if it[2] == kind]
            if not group:
                continue
            try:
                images = torch.stack([it[0] for it in group])
                have = [it[1] is not None for it in group]
                idx = None
                if all(have):
                    idx = np.stack([it[1] for it in group])
                elif any(have):                               # draw the missing ones so the batch stays one call
                    drawn = self.solver.draw_indices(len(group))
                    idx = np.stack([it[1] if it[1] is not None else drawn[i] for i, it in enumerate(group)])
                if self.graph:
                    pad = self.max_batch - images.shape[0]
                    if idx is None:
                        idx = self.solver.draw_indices(len(group))
                    if pad > 0:
                        images = torch.cat([images, images[-1:].expand(pad, -1, -1, -1)])
                        idx = np.concatenate([idx, np.repeat(idx[-1:], pad, axis=0)])
                    res = self.solver.solve(images, indices=idx, want_images=True, prescrambled=kind, graph=True,
                                            **({"count_rows": len(group)} if pad > 0 else {}))   # padding stays out of the totals
                else:
                    res = self.solver.solve(images, indices=idx, want_images=True, prescrambled=kind)
                ok, matches = res.puzzle_correct.tolist(), res.patch_matches.tolist()
                truth, pred = res.indices.tolist(), res.pred.tolist()
                for i, it in enumerate(group):
                    it[3].set_result({"puzzle_correct": int(ok[i]), "patch_matches": int(matches[i]), "total_patches": n,
                                      "patch_accuracy": matches[i] / n, "indices": truth[i], "predicted_order": pred[i],
                                      "solution_image": res.reconstructed[i], "scrambled_image": res.scrambled[i]})
            except Exception as e:                            # the API answers every request, failed or not (app.py:346-348)
                for it in group:
                    if not it[3].done():
                        it[3].set_exception(e)
        self.batches_run += 1

    def _loop(self) -> None:
        dev = getattr(self.solver, "device", None)
        if dev is not None and dev.type == "cuda":
            torch.cuda.set_device(dev)          # a new thread starts on device 0 whatever the model's device
        while not self._stop.is_set():
            items = self._collect()
            if items is None:
                break
            self._run(items)
        while True:                                            # fail whatever is still queued
            try:
                it = self._q.get_nowait()
            except queue.Empty:
                break
            if it is not None and not it[3].done():
                it[3].set_exception(RuntimeError("MicroBatcher closed"))
