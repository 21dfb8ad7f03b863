#!/usr/bin/env python
"""Does running the tensor-bound and the MUFU-bound halves of the forward concurrently pay on B200?

Two engines (two handles, two workspaces) score half a batch each on two CUDA streams, so one half's GEMMs can run under
the other half's scan / conv.  Compared with one engine scoring the whole batch on one stream.  Reports ms per 256
windows, SM clock and power for both schedules.  

    python tools/overlap_probe.py [--model l32] [--batch 256] [--steps 8]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sample_smi(stop, out):
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            out.append((float(r[0]), float(r[1])))
        except Exception:
            pass
        time.sleep(0.2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="l32")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--offset", type=int, default=0, help="stagger: stream B starts after this many ms of stream A's step (sleep on host)")
    args = ap.parse_args()
    import torch
    from plantcaduceus_b200 import preset, random_init_state_dict
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    dev = torch.device("cuda:0")
    cfg = preset(args.model)
    sd = random_init_state_dict(cfg, seed=0)
    B, L = args.batch, 512
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(3, 7, (B, L), generator=g).to(torch.uint8)
    ids[:, 255] = 1
    ids = ids.to(dev)
    pos = torch.full((B, 1), 255, dtype=torch.int32, device=dev)
    res = {}

    def timed(fn, label):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        stop, samples = threading.Event(), []
        th = threading.Thread(target=sample_smi, args=(stop, samples), daemon=True)
        th.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps * 1e3
        stop.set()
        th.join()
        sm = sorted(s[0] for s in samples)
        pw = sorted(s[1] for s in samples)
        res[label] = {"ms_per_batch": dt, "windows_per_s": B / dt * 1e3, "sm_mhz_median": sm[len(sm) // 2] if sm else None,
                      "power_w_median": pw[len(pw) // 2] if pw else None}
        print(label, res[label], flush=True)

    one = CaduceusForMaskedLM.from_pretrained(sd, config=cfg, torch_dtype=torch.bfloat16).to(dev)
    timed(lambda: one.score_masked(ids, pos, check_ids=False), "one_stream_full_batch")
    timed(lambda: (one.score_masked(ids[:B // 2], pos[:B // 2], check_ids=False),
                   one.score_masked(ids[B // 2:], pos[B // 2:], check_ids=False)), "one_stream_two_half_batches")
    two = CaduceusForMaskedLM.from_pretrained(sd, config=cfg, torch_dtype=torch.bfloat16).to(dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    outs = [None, None]

    def both():
        with torch.cuda.stream(s1):
            outs[0] = one.score_masked(ids[:B // 2], pos[:B // 2], check_ids=False)
        with torch.cuda.stream(s2):
            outs[1] = two.score_masked(ids[B // 2:], pos[B // 2:], check_ids=False)
    # one warm-up with a stagger so the two streams run out of phase (A's GEMMs under B's scan) rather than in lock step
    with torch.cuda.stream(s2):
        two.score_masked(ids[: B // 4], pos[: B // 4], check_ids=False)
    timed(both, "two_streams_half_batch_each")
    full = one.score_masked(ids, pos)
    both()
    torch.cuda.synchronize()
    res["bitwise_equal"] = bool(torch.equal(torch.cat(outs), full))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_overlap.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
