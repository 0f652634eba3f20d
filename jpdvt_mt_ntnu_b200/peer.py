"""NVLink / NVSwitch peer-memory plumbing for the data-parallel optimizer step (csrc/peer_optim.cu).

Every rank allocates ONE symmetric block (torch symmetric memory: CUDA VMM handles exchanged over the process group, plus
the NVSwitch multicast object when the fabric offers one) laid out as

    [ flat fp32 gradients | flat bf16 operand copy | fp32 parameters | Adam m | Adam v | EMA | flag pad ]

and hands the kernel the addresses of every rank's block as mapped into THIS process.  PyTorch is plumbing here: it owns
the allocation and the rendezvous; no collective runs on the data path - the reduce-scatter, the optimizer and the
all-gather are one kernel (`jpdvt_adamw_ema_peer`) that reads and writes peer memory directly.

Replaces the gradient exchange of DistributedDataParallel in the reference trainer (image_model/train_JPDVT.py:231, 370).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import MAX_F32_RANGES, MAX_PEERS, PeerStep


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(chunk, begin, end): equal contiguous slices of the flat parameter space, a multiple of 2048 parameters each (the
    tile the bulk-copy kernel moves: 8 KB of fp32 gradients per peer, 4 KB of bf16 results); buffers are padded to
    world * chunk."""
    chunk = -(-total // (world * 2048)) * 2048
    return chunk, rank * chunk, (rank + 1) * chunk


def merge_ranges(ranges, limit: int = MAX_F32_RANGES):
    """Sorted union of half-open index ranges (touching ranges are joined); raises if more than `limit` remain."""
    merged = []
    for b, e in sorted((int(b), int(e)) for b, e in ranges if e > b):
        if merged and b <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], e)
        else:
            merged.append([b, e])
    if len(merged) > limit:
        raise _lib.JpdvtError(f"{len(merged)} fp32 ranges, the kernel takes {limit}")
    return [tuple(r) for r in merged]


def available(group=None) -> Tuple[bool, str]:
    """Can this process group use the peer-memory step?  (one node, 2..8 ranks, NCCL group, symmetric memory present)"""
    if not dist.is_initialized():
        return False, "torch.distributed is not initialised"
    world = dist.get_world_size(group)
    if world < 2 or world > MAX_PEERS:
        return False, f"world size {world} outside 2..{MAX_PEERS}"
    if dist.get_backend(group) != "nccl":
        return False, f"backend {dist.get_backend(group)} (needs CUDA ranks)"
    if int(os.environ.get("LOCAL_WORLD_SIZE", world)) != world:
        return False, "ranks span several nodes"
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
    except Exception as e:  # noqa: BLE001
        return False, f"torch symmetric memory unavailable: {e}"
    return True, ""


class PeerExchange:
    """The symmetric block of one rank + the mapped addresses of everyone else's."""

    FLAG_WORDS = 2 * MAX_PEERS + 2          # barrier A flags, barrier B flags, local CTA counter, status word

    def __init__(self, total: int, device: torch.device, group=None, multicast: Optional[bool] = None, timeout_ms: int = 20000):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.total = total
        self.chunk, self.begin, self.end = shard_bounds(total, self.world, self.rank)
        self.padded = self.chunk * self.world
        self.device = device
        n = self.padded
        # section offsets in bytes (all multiples of 16): gradients, bf16 operands, then the fp32 state (parameters, Adam
        # moments, EMA).  The state is mapped too so that any rank can READ any owner's slice one-sidedly (checkpoints).
        self._off = {"grads": 0, "weights": 4 * n, "p": 6 * n, "m": 10 * n, "v": 14 * n, "ema": 18 * n, "flags": 22 * n}
        nbytes = 22 * n + 4 * 64
        with torch.cuda.device(device):
            self.block = symm.empty(nbytes, dtype=torch.uint8, device=device)
            self.block.zero_()
            torch.cuda.synchronize(device)
            self.handle = symm.rendezvous(self.block, self.group.group_name)
        sec = lambda name, width, dtype: self.block[self._off[name]:self._off[name] + width * n].view(dtype)
        self.grads = sec("grads", 4, torch.float32)
        self.weights_bf16 = sec("weights", 2, torch.bfloat16)
        self.p, self.m, self.v, self.ema = (sec(k, 4, torch.float32) for k in ("p", "m", "v", "ema"))
        self.flags = self.block[self._off["flags"]:self._off["flags"] + 4 * 64].view(torch.int32)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.block.data_ptr():
            raise _lib.JpdvtError("symmetric-memory rendezvous returned an unexpected address table")
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        env = os.environ.get("JPDVT_PEER_MULTICAST")
        if multicast is None:           # measured slower than bulk peer copies on 2 GPUs: opt-in
            multicast = env == "1"
        self.multicast = bool(multicast and mc)
        s = PeerStep()
        s.world, s.rank, s.shard_begin, s.shard_end = self.world, self.rank, self.begin, self.end
        s.epoch, s.timeout_ms = 0, timeout_ms
        for q in range(self.world):
            s.grads[q] = ptrs[q] + self._off["grads"]
            s.weights_bf16[q] = ptrs[q] + self._off["weights"]
            s.params[q] = ptrs[q] + self._off["p"]
            s.signals[q] = ptrs[q] + self._off["flags"]
        s.grads_mc = (mc + self._off["grads"]) if self.multicast else None
        s.weights_mc = (mc + self._off["weights"]) if self.multicast else None
        s.params_mc = (mc + self._off["p"]) if self.multicast else None
        s.n_f32_ranges, s.reserved = 0, 0
        s.local_sync = self.flags.data_ptr() + 4 * (2 * MAX_PEERS)
        s.status = self.flags.data_ptr() + 4 * (2 * MAX_PEERS + 1)
        # the barrier token lives on the device (the kernel reads *epoch_dev + 1 and stores it back), so a CUDA-graph replay
        # of the step needs no changing kernel parameter; every rank starts from 0 and steps in lockstep
        s.epoch_dev = self.flags.data_ptr() + 4 * (2 * MAX_PEERS + 2)
        self.struct = s
        self.epoch = 0
        self.handle.barrier()                       # every rank's block is zeroed and mapped before anyone's first step

    def set_f32_ranges(self, ranges) -> None:
        """Index ranges of the flat parameter space whose fp32 values every rank needs after a step (the parameters the
        kernels read in fp32: biases, timestep MLP, position head) - merged, at most MAX_F32_RANGES."""
        merged = merge_ranges(ranges)
        self.struct.n_f32_ranges = len(merged)
        for k, (b, e) in enumerate(merged):
            self.struct.f32_ranges[2 * k], self.struct.f32_ranges[2 * k + 1] = b, e
        self.f32_ranges = merged

    def pull(self, names=("p", "m", "v", "ema")) -> None:
        """One-sided gather: copy every other owner's slice of the named fp32 state buffers into this rank's copy (peer
        reads over NVLink, stream ordered, no participation of the other ranks).  Between two steps the owners' slices
        are consistent: no rank can finish step k+1's kernel before THIS rank has launched its own (barrier A)."""
        with torch.cuda.device(self.device):
            for name in names:
                local = getattr(self, name)
                for q in range(self.world):
                    if q == self.rank:
                        continue
                    remote = self.handle.get_buffer(q, (self.chunk,), torch.float32, (self._off[name] // 4) + q * self.chunk)
                    local[q * self.chunk:(q + 1) * self.chunk].copy_(remote)

    def next_epoch(self) -> PeerStep:
        """Barrier token of the next step: strictly increasing, identical on every rank (all ranks step in lockstep)."""
        self.epoch += 1
        self.struct.epoch = self.epoch
        return self.struct

    def check(self) -> None:
        """Host-side look at the status word (one small device->host copy; call it at logging / checkpoint cadence)."""
        st = int(self.flags[2 * MAX_PEERS + 1].item())
        if st != 0:
            which = {1: "waiting for the peers' gradients", 2: "waiting for the peers' weight stores"}.get(st, str(st))
            raise _lib.JpdvtError(f"peer-memory optimizer step timed out {which}: a rank died or skipped a step")

    def describe(self) -> str:
        how = ("multimem.ld_reduce / multimem.st through the NVSwitch multicast object" if self.multicast
               else "per-thread peer loads / stores over NVLink" if os.environ.get("JPDVT_PEER_VARIANT", "")[:1] == "t"
               else "cp.async.bulk pulls of the peers' gradient tiles and pushes of the bf16 tiles over NVLink")
        return (f"fused reduce-scatter + AdamW/EMA on 1/{self.world} of the state + bf16 all-gather in one kernel per rank "
                f"({how}; {self.total * 4 / 1e6:.0f} MB of fp32 gradients, no NCCL call on the data path)")
