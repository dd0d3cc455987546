"""Product-sumcheck kernel variants side by side (development aid, run under gpurun): CTA size, shared-memory lazy
accumulators and the number of folds moved to the FP64 pipe.  The variants are chosen by environment variables the
library reads once, so every point runs in its own process.  Prints one JSON line per point; every point must
produce the same messages (sha256 of the canonical bytes) as the baseline point."""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(v):
    sys.path.insert(0, ROOT)
    import gkr_b200
    from gkr_b200 import synthetic as syn
    pv = gkr_b200.Prover(0)
    tabs = [pv.dev_table_synth(1, syn.TABLE_STREAM + t, 1 << v) for t in range(3)]
    pv.sumcheck_prod_raw(tabs, v)                        # warm-up
    times = []
    for _ in range(3):
        pv.sync()
        t0 = time.perf_counter()
        res = pv.sumcheck_prod_raw(tabs, v)
        times.append((time.perf_counter() - t0) * 1e3)
    pv.profile(1)
    pv.sumcheck_prod_raw(tabs, v)
    prof = pv.profile(0)
    h = hashlib.sha256()
    for part in res:
        h.update(part.tobytes())
    algo = 32 * 3 * (4 * (1 << v) - 6)
    out = {"vars": v, "ms": round(min(times), 3), "tbs_whole": round(algo / min(times) / 1e9, 3),
           "prod3_first_ms": round(prof["prod3_round"]["ms"], 3), "prod3_fused_ms": round(prof["prod3_round_fused"]["ms"], 3),
           "prod3_fused_tbs": round(prof["prod3_round_fused"]["algo_bytes"] / max(prof["prod3_round_fused"]["ms"], 1e-9) / 1e9, 3),
           "prod3_first_tbs": round(prof["prod3_round"]["algo_bytes"] / max(prof["prod3_round"]["ms"], 1e-9) / 1e9, 3),
           "sha": h.hexdigest()[:16]}
    print(json.dumps(out))


def main():
    v = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    points = []
    if len(sys.argv) > 2 and sys.argv[2] == "exact":
        # the exact-product evaluation (k_prod3_round_x, fr_wide3.cuh) next to the default kernel: GKR_P3_EXACT = 1 + KA
        # (KA bit 0 / 1 = Karatsuba in the first / second stage), CTA size of the fused rounds
        points.append({})
        for threads in (512, 384):
            for mode in (1, 2, 3, 4):
                points.append({"GKR_P3_EXACT": str(mode), "GKR_P3X_THREADS": str(threads)})
    else:
        for threads in (256, 384, 448, 512):
            for sacc in (0, 1):
                for nf in (0,):
                    points.append({"GKR_P3_THREADS": str(threads), "GKR_P3_SACC": str(sacc), "GKR_F64_FOLDS": str(nf)})
    base = None
    for env in points:
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, __file__, "--child", str(v)], env=e, capture_output=True, text=True, timeout=600)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        try:
            d = json.loads(line)
        except Exception:
            print(json.dumps({"env": env, "error": (r.stderr or r.stdout)[-400:]}))
            continue
        if base is None:
            base = d["sha"]
        d["same_as_baseline"] = d["sha"] == base
        d["env"] = env
        print(json.dumps(d), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        main()
