#!/usr/bin/env python3
"""Script-level drop-in run: the UNMODIFIED reference scripts `inference_ddp.py` and `train_JPDVT.py`
(/root/reference/image_model, staged by tools/stage_reference.sh under the git-ignored baseline/_ref/) executed on this
library through nothing but an import-path switch:

    PYTHONSAFEPATH=1  PYTHONPATH=<repo>/dropin:<standins>:<reference>/image_model  torchrun ... <script>

PYTHONSAFEPATH keeps the interpreter from putting the script's own directory first on sys.path, so `from models import
DiT_models` / `from diffusion import create_diffusion` (inference_ddp.py:40-41, train_JPDVT.py:24-26) resolve to
dropin/ while `from datasets import MET, TEXMET` still finds the reference's own file.  Stand-ins cover the two absent
third-party imports the scripts never use (diffusers.AutoencoderKL, matplotlib.pyplot).

What this harness fabricates, because the scripts hard-code it: the data directory, the checkpoint path and the output
directories of inference_ddp.py (`/cluster/home/muhamhz/...`, lines 48-62) and an ImageFolder tree for train_JPDVT.py.
Images are synthetic JPEGs; the checkpoint holds seeded random weights in the reference's `{"model": state_dict}` form.

Writes logs under gpurun_out/dropin/ and exits non-zero if a script failed, logged a per-image error, produced no CSV
rows / checkpoint, or if the reference's own models.py was imported instead of the drop-in.
"""
import csv
import glob
import os
import shutil
import subprocess
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref", "image_model")
OUT = os.path.join(ROOT, "gpurun_out", "dropin")
CLUSTER = "/cluster/home/muhamhz"
N_INFER = int(os.environ.get("N_INFER", "12"))
N_TRAIN = int(os.environ.get("N_TRAIN", "32"))
NPROC = int(os.environ.get("NPROC", "1"))


def synth_image(path, seed, size=320):
    """A smooth random picture with structure at the piece scale (so JPEG keeps it) - content is irrelevant here."""
    rng = np.random.RandomState(seed)
    low = rng.rand(10, 10, 3)
    img = np.asarray(Image.fromarray((low * 255).astype(np.uint8)).resize((size, size), Image.BICUBIC), dtype=np.float32)
    img += rng.randn(size, size, 3) * 12
    Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(path, quality=92)


def env():
    e = dict(os.environ)
    e["PYTHONSAFEPATH"] = "1"
    e["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "dropin"), os.path.join(ROOT, "tools", "standins"), REF])
    e["JPDVT_DROPIN_VERBOSE"] = "1"
    e["WANDB_MODE"] = "disabled"
    e.setdefault("OMP_NUM_THREADS", "4")
    return e


def torchrun(script, args, log, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={NPROC}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, *args]
    with open(log, "w") as f:
        f.write("$ PYTHONSAFEPATH=1 PYTHONPATH=dropin:tools/standins:<reference>/image_model " + " ".join(cmd[1:]) + "\n")
        f.flush()
        r = subprocess.run(cmd, cwd=REF, env=env(), stdout=f, stderr=subprocess.STDOUT, timeout=1500)
    return r.returncode


def main():
    if not os.path.isdir(REF):
        sys.exit("baseline/_ref/image_model is not staged (run tools/stage_reference.sh where /root/reference exists)")
    os.makedirs(OUT, exist_ok=True)
    fails = []

    # ---- inference_ddp.py -------------------------------------------------------------------------------------------
    data = f"{CLUSTER}/data/imagenet/test"
    shutil.rmtree(f"{CLUSTER}/JPDVT/image_model/logs_fresh", ignore_errors=True)
    os.makedirs(data, exist_ok=True)
    for i in range(N_INFER):
        synth_image(f"{data}/synthetic_{i:04d}.JPEG", i)
    from jpdvt_mt_ntnu_b200.models import DiT_models
    from jpdvt_mt_ntnu_b200.weights import seeded_state
    model = DiT_models["JPDVT"](input_size=192)
    ck = f"{CLUSTER}/JPDVT/image_model/models/3x3_Full/2850000.pt"
    os.makedirs(os.path.dirname(ck), exist_ok=True)
    torch.save({"model": seeded_state(model.state_dict(), seed=1234)}, ck)
    rc = torchrun("inference_ddp.py", [], f"{OUT}/inference_ddp.log", 29611)
    logs = f"{CLUSTER}/JPDVT/image_model/logs_fresh"
    rows = []
    if os.path.exists(f"{logs}/fresh_inference_progress.csv"):
        rows = list(csv.DictReader(open(f"{logs}/fresh_inference_progress.csv")))
        shutil.copy(f"{logs}/fresh_inference_progress.csv", f"{OUT}/inference_ddp_progress.csv")
    errs = open(f"{logs}/inference_errors.txt").read() if os.path.exists(f"{logs}/inference_errors.txt") else ""
    pngs = glob.glob(f"{CLUSTER}/JPDVT/image_model/inference_fresh/Grid3/*_combined_*.png")
    text = open(f"{OUT}/inference_ddp.log").read()
    if rc != 0:
        fails.append(f"inference_ddp.py exit code {rc}")
    if len(rows) != N_INFER:
        fails.append(f"inference_ddp.py wrote {len(rows)} CSV rows, expected {N_INFER}")
    if errs.strip():
        fails.append("inference_ddp.py logged per-image errors:\n" + errs[:2000])
    if len(pngs) < N_INFER:
        fails.append(f"only {len(pngs)} combined PNGs")
    if "[jpdvt-dropin] models ->" not in text or "[jpdvt-dropin] diffusion ->" not in text:
        fails.append("inference_ddp.py did not import the drop-in modules")

    # ---- train_JPDVT.py (ImageFolder, 288 px: the one imagenet setting whose final validation type-checks) -----------
    tdata = os.path.join(OUT, "imagefolder")
    shutil.rmtree(tdata, ignore_errors=True)
    for split, n in (("train", N_TRAIN), ("val", 4)):
        for i in range(n):
            d = f"{tdata}/{split}/class{i % 2}"
            os.makedirs(d, exist_ok=True)
            synth_image(f"{d}/img_{i:04d}.JPEG", 1000 + i)
    results = os.path.join(OUT, "train_results")
    shutil.rmtree(results, ignore_errors=True)
    bs = 8 * NPROC
    rc = torchrun("train_JPDVT.py", ["--data-path", f"{tdata}/train", "--dataset", "imagenet", "--image-size", "288",
                                     "--epochs", "2", "--global-batch-size", str(bs), "--num-workers", "2",
                                     "--log-every", "1", "--ckpt-every", "1000", "--results-dir", results,
                                     "--disable-wandb"], f"{OUT}/train_JPDVT.log", 29612)
    text = open(f"{OUT}/train_JPDVT.log").read()
    finals = glob.glob(f"{results}/*/checkpoints/final_*.pt")
    if rc != 0:
        fails.append(f"train_JPDVT.py exit code {rc}")
    if not finals:
        fails.append("train_JPDVT.py wrote no final checkpoint")
    else:
        ckpt = torch.load(finals[0], weights_only=False)
        keys = sorted(ckpt.keys())
        with open(f"{OUT}/train_JPDVT_checkpoint.txt", "w") as f:
            f.write(f"{os.path.basename(finals[0])}: keys {keys}, train_steps {ckpt['train_steps']}, "
                    f"{len(ckpt['model'])} model tensors, {len(ckpt['opt']['state'])} optimizer states\n")
        if ckpt["train_steps"] != 2 * (N_TRAIN // bs):
            fails.append(f"train_steps {ckpt['train_steps']}")
    if "Train Loss" not in text or "Done!" not in text:
        fails.append("train_JPDVT.py log lacks 'Train Loss' / 'Done!'")
    if "[jpdvt-dropin] models ->" not in text:
        fails.append("train_JPDVT.py did not import the drop-in modules")
    for d in (results, tdata):
        shutil.rmtree(d, ignore_errors=True)          # checkpoints are 2 GB: keep gpurun_out small

    with open(f"{OUT}/summary.txt", "w") as f:
        f.write(("FAILED\n" + "\n".join(fails) + "\n") if fails else
                f"OK: inference_ddp.py {len(rows)} puzzles, train_JPDVT.py {2 * (N_TRAIN // bs)} steps, nproc {NPROC}\n")
    print(open(f"{OUT}/summary.txt").read())
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
