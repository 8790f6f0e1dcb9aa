#!/usr/bin/env python
"""Frozen-input annihilation microbenchmark: n spawn records with keys drawn from D
distinct addresses (duplicate ratio D/n), Float64 values, resident in HBM; the three methods of
rimu_annihilate_device are timed with CUDA events and reported as GB/s of the bytes that MUST move
(n*E read + U*E written, E = 16).  Output: one JSON line per (n, D/n, method) + a markdown table.

    python bench_annihilate.py [--log2n 24] > profiles/annihilation_methods.md
"""
import argparse
import ctypes as C
import json
import sys

import numpy as np
import torch

import rimu_b200 as R
from rimu_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, nargs="+", default=[22, 24, 26])
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6650.0)
    at = R.AddressType(_lib.ADDR_BOSE, (20,), 20)
    dev = torch.device("cuda", 0)
    rows = []
    names = {0: "hash (HBM table)", 1: "sort (radix + segmented reduce)", 2: "partition (bucket streams + smem merge)"}
    for lg in a.log2n:
        n = 1 << lg
        ctx = R.get_context(1)
        ctx.resize_table(4 * n)
        for ratio in (1.0, 0.25, 1 / 64):
            D = max(1, int(n * ratio))
            g = torch.Generator(device=dev).manual_seed(42)
            pool = torch.randint(1, 2 ** 39, (D,), dtype=torch.int64, device=dev, generator=g)
            keys = pool[torch.randint(0, D, (n,), device=dev, generator=g)].contiguous()
            vals = (torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 - 1).contiguous()
            ref = None
            for method in (0, 1, 2):
                v = R.GPUDVec(style=R.IsDeterministic(), address_type=at, capacity=n + 1024)
                best = 1e30
                for rep in range(a.reps + 1):
                    ms = C.c_float()
                    _lib.check(_lib.lib().rimu_annihilate_device(v.handle, C.c_void_p(keys.data_ptr()), C.c_void_p(vals.data_ptr()), n, method, C.byref(ms)))
                    if rep:
                        best = min(best, ms.value)
                U = len(v)
                s1 = v.norm(1)
                if ref is None:
                    ref = (U, s1)
                assert U == ref[0] and abs(s1 - ref[1]) <= 1e-9 * ref[1], (method, U, ref)
                gbs = (n * 16 + U * 16) / (best * 1e-3) / 1e9
                row = {"n": n, "distinct_ratio": ratio, "method": names[method], "ms": best, "unique": U, "gbs": gbs, "frac_of_hbm_peak": gbs / peak}
                rows.append(row)
                print(json.dumps(row), file=sys.stderr, flush=True)
                del v
    print("| records n | distinct/n | method | ms | unique out | algorithmic GB/s | of measured HBM peak |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| 2^{int(np.log2(r['n']))} | {r['distinct_ratio']:.4g} | {r['method']} | {r['ms']:.3f} | {r['unique']} | {r['gbs']:.0f} | {100 * r['frac_of_hbm_peak']:.1f} % |")


if __name__ == "__main__":
    main()
