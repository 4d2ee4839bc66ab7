#!/usr/bin/env python
"""Run every GPU parity check and print ALL results (no early stop) -- the blind-debugging entry point for gpurun.

  python tests/run_gpu_checks.py [--precise] [--only substr] [--out gpurun_out/checks.json]
"""
import argparse
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precise", action="store_true", help="3xTF32 tensor-core products (isolates indexing bugs)")
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "checks.json"))
    args = ap.parse_args()
    import torch
    from tests import checks
    from easevoice_trainer_b200 import lib
    L = lib.init()
    L.evk_set_precise(1 if args.precise else 0)
    checks.PRECISE_MODE[0] = bool(args.precise)
    results, failed = [], 0
    names = checks.NAMES
    for nm, fn in zip(names, checks.ALL):
        if args.only and args.only not in nm:
            continue
        t0 = time.time()
        try:
            rows = fn()
            torch.cuda.synchronize()
        except Exception:
            tb = traceback.format_exc()
            print(f"!! {nm}: EXCEPTION\n{tb}", flush=True)
            results.append(dict(group=nm, name="EXCEPTION", err=None, tol=None, ok=False, tb=tb))
            failed += 1
            try:
                torch.cuda.synchronize()
            except Exception as e:      # sticky CUDA error: nothing else can run in this process
                print("CUDA context is dead:", e, flush=True)
                break
            continue
        for name, err, tol in rows:
            ok = (err == err) and err <= tol
            failed += (not ok)
            results.append(dict(group=nm, name=name, err=err, tol=tol, ok=ok))
            print(f"{'ok  ' if ok else 'FAIL'} {name:72s} err={err:.3e} tol={tol:.1e}", flush=True)
        print(f"-- {nm}: {time.time() - t0:.1f}s", flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(dict(precise=args.precise, failed=failed, results=results), f, indent=1)
    if checks.REPORT:
        with open(os.path.join(os.path.dirname(args.out), "parity_table.json"), "w") as f:
            json.dump(checks.REPORT, f, indent=1)
    print(f"TOTAL {len(results)} checks, {failed} failed (precise={args.precise})")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
