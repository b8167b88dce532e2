"""Developer probe (run under gpurun): FFMA peak, spot-check against the direct scan, per-mode timing at growing sizes."""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hvs = importlib.import_module("project---hybrid-vector-search-queries_b200")

def run(n, m, ncat, modes, types=(0, 1, 2, 3), check_n=0, reps=2):
    d = hvs.gen_data(n, 3, ncat=ncat)
    q = hvs.gen_queries(m, 4, ncat=ncat, types=types)
    res = {}
    for name, mode in modes:
        with hvs.Engine(mode=mode) as e:
            t0 = time.time(); e.index_build(d); tb = time.time() - t0
            best = None
            for r in range(reps):
                ids = e.solve(q); st = e.stats()
                if best is None or st["ms_solve_device"] < best["ms_solve_device"]: best = st
            st = best
            res[name] = ids
            tf = 200.0 * st["pairs"] / (st["ms_solve_device"] * 1e-3) / 1e12
            print(f"n={n} m={m} ncat={ncat} types={types} mode={name}: solve_dev={st['ms_solve_device']:.2f}ms wall={st['ms_solve_wall']:.2f}ms "
                  f"qps={m / (st['ms_solve_device'] * 1e-3):.0f} algTF={tf:.2f} plan={st['ms_plan']:.2f} direct={st['ms_direct']:.2f} tile={st['ms_tile']:.2f} "
                  f"fin={st['ms_finalize']:.2f} index={st['ms_index_build']:.1f}ms(wall {tb:.2f}s) ndirect={st['n_direct']} ntile={st['n_tile']} items={st['n_items_ffma']}+{st['n_items_tensor']} "
                  f"fallback={st['n_fallback']} pairs={st['pairs']:.3e} computed={st['pairs_computed']:.3e}", flush=True)
    names = list(res)
    for a in names[1:]:
        same = np.array_equal(res[names[0]], res[a])
        print(f"   ids {names[0]} == {a}: {same}")
    if check_n:
        # spot check against the direct scan (the reference's arithmetic by construction; tests/ pin it to the oracle):
        # re-scored distances must be the same fp32 numbers, position by position
        pick = np.linspace(0, m - 1, check_n).astype(int)
        with hvs.Engine(mode=hvs.MODE_DIRECT) as e:
            e.index_build(d)
            ref_ids = e.solve(q[pick])
            ref_dist = e.rescore(q[pick], ref_ids)
            for a in names:
                dist = e.rescore(q[pick], res[a][pick])
                same = np.array_equal(dist.view(np.uint32), ref_dist.view(np.uint32))
                print(f"   direct-scan check {a}: queries={check_n} ok={same} id_rows_differ={int((res[a][pick] != ref_ids).any(axis=1).sum())}")

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ALL = {"direct": hvs.MODE_DIRECT, "exact": hvs.MODE_EXACT, "auto": hvs.MODE_AUTO}
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["direct", "exact", "auto"]
    M = [(k, ALL[k]) for k in names]
    FAST = [(k, ALL[k]) for k in names if k != "direct"]
    if which in ("all", "peak"):
        with hvs.Engine() as e:
            print("ffma peak TF/s, MHz:", e.measure_ffma_peak(5), flush=True)
    if which in ("all", "small"):
        run(10_000, 100, 10, M, check_n=100)
        run(200_000, 1024, 10, M, check_n=32)
    if which in ("all", "mid"):
        run(1_000_000, 4096, 100, M, check_n=16)
        run(2_000_000, 2048, 10, FAST, types=(0,), check_n=4)
    if which in ("all", "big"):
        run(10_000_000, 40_000, 100, FAST, check_n=8, reps=2)
        run(10_000_000, 40_000, 100, FAST, types=(0,), check_n=4, reps=2)
