#!/usr/bin/env python
"""Achieved HBM GB/s of the bandwidth-bound kernels from an ncu launch list (gpu__time_duration.sum) of ONE learn()
iteration, against their ALGORITHMIC bytes (SURVEY 8d / DESIGN 4).
usage: python profiles/hbm_kernels_table.py <launches.csv> <envs> <obs_dim> [peak_gbs]"""
import collections, csv, re, sys

path, N, O = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
peak = float(sys.argv[4]) if len(sys.argv) > 4 else 6555.5
T, A, M, E = 24, 12, 4, 5
B = T * N
mb = B // M
dims_a, dims_c = [O, 512, 256, 128, A], [O, 512, 256, 128, 1]
P = sum(a * b + b for d in (dims_a, dims_c) for a, b in zip(d[:-1], d[1:])) + A
row = 4 * O + 12 * A + 16
ALG = {  # kernel-name fragment -> (algorithmic bytes per launch, description)
    "gae_norm": (17 * T * N + 4 * N, "GAE reverse scan + advantage normalisation"),
    "store_transition": (2 * N * (4 * O + 17) , "rollout-buffer row write (obs + rewards + dones; act() wrote the rest in place)"),
    "permute_gather": (B * (2 * row + 8), "permuted gather of the whole rollout (once per update)"),
    "ppo_loss": ((20 * A + 24) * mb, "PPO objective forward + backward"),
    "gaussian_act": (N * (12 * A + 4 + 4 * A), "sample + log-prob (reads mu, z; writes actions, sigma, logp)"),
    "allreduce_clip_adam": (32 * P, "grad-norm + clip + KL-LR + Adam (one launch)"),
    "grad_norm_lr": (4 * P, "grad-norm + clip + KL-LR"),
    "adam_step": (28 * P, "Adam"),
}
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
agg = collections.defaultdict(list)
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] in ("ns", "nsecond") else v
    for frag in ALG:
        if frag in r[ki]:
            agg[frag].append(v)
print(f"# envs {N}, obs {O}, B = {B}, minibatch {mb}, P = {P}; peak {peak} GB/s (MEASURED_PEAKS.json hbm_gbs); ncu per-launch "
      f"durations (cold caches, serialised)")
print(f"{'kernel':22s} {'launches':>8s} {'alg. MB':>9s} {'median us':>10s} {'GB/s':>8s} {'frac':>6s}  what")
for frag, ts in agg.items():
    ts.sort()
    med = ts[len(ts) // 2]
    nbytes, what = ALG[frag]
    gbs = nbytes / med / 1e3
    print(f"{frag:22s} {len(ts):8d} {nbytes/1e6:9.2f} {med:10.2f} {gbs:8.1f} {gbs/peak:6.3f}  {what}")
