"""Full-depth oracle outputs for the published sizes (l24 x 24, l28 x 28, l32 x 32 layers), generated ONCE in the
build container (the CPU oracle needs minutes per size) and committed as fixtures:

    python tests/golden/make_fulldepth_golden.py [l24 l28 l32 l32g]

Each ``fulldepth_<tag>.npz`` holds, for the seeded inputs below (weights = random_init_state_dict(preset, seed=0), rebuilt by
the tests from the same seed -- nothing but the seed travels):
    ids          int64 [B, 512]   (a/c/g/t uniform, one N, index 255 masked)
    logits_f32   float32 [B, 512, 8]   oracle.caduceus_forward(dtype=float32)
    logits_bf16  float32 [B, 512, 8]   oracle.caduceus_forward(dtype=bfloat16)  -- the reference algorithm's own bf16 drift
``l32g`` is l32 with NON-TRIVIAL norm gains (log-uniform in [0.25, 4], seeded): trained checkpoints have large RMSNorm
weights, which is where folding the norm weight into in_proj's columns could round differently.
The oracle is the checker; the product never reads these files.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import caduceus_oracle as O  # noqa: E402
from plantcaduceus_b200 import preset, random_init_state_dict  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
B, L, IDX = 8, 512, 255
SIZES = {"l24": "PlantCaduceus_l24", "l28": "PlantCaduceus_l28", "l32": "PlantCaduceus_l32", "l32g": "PlantCaduceus_l32"}


def make_inputs(tag):
    g = torch.Generator().manual_seed(1000 + len(tag) + sum(map(ord, tag)))
    ids = torch.randint(3, 7, (B, L), generator=g)
    ids[:, IDX] = 1
    ids[0, 7] = 2          # an N
    return ids


def make_weights(tag):
    cfg = preset(SIZES[tag])
    sd = random_init_state_dict(cfg, seed=0)
    if tag.endswith("g"):
        g = torch.Generator().manual_seed(77)
        for k in list(sd):
            if k.endswith("norm.weight") or k.endswith("norm_f.weight"):
                gain = torch.exp((torch.rand(sd[k].shape, generator=g) * 2 - 1) * np.log(4.0))
                sign = torch.where(torch.rand(sd[k].shape, generator=g) < 0.05, -1.0, 1.0)   # a few negative gains too
                sd[k] = (gain * sign).float()
    return cfg, sd


def main():
    tags = sys.argv[1:] or list(SIZES)
    torch.set_num_threads(os.cpu_count() or 1)
    for tag in tags:
        cfg, sd = make_weights(tag)
        ids = make_inputs(tag)
        out = {"ids": ids.numpy()}
        for name, dt in (("logits_f32", torch.float32), ("logits_bf16", torch.bfloat16)):
            t0 = time.time()
            with torch.inference_mode():
                logits, _ = O.caduceus_forward(sd, cfg, ids, dtype=dt)
            out[name] = logits.float().numpy()
            print(f"{tag} {name}: {time.time() - t0:.0f} s", flush=True)
        np.savez_compressed(os.path.join(HERE, f"fulldepth_{tag}.npz"), **out)


if __name__ == "__main__":
    main()
