"""Step-by-step comparison with the oracle on the default shape; stops at the first mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from __graft_entry__ import load_package
fsb = load_package()
from oracle import oracle as orc
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
p = fsb.Params.default()
w = orc.World(p, seed=seed, rep=0); w.enable_trace(); w.initialise()
hz = np.asarray(w.horizon)
with fsb.Engine(p, 2) as eng:
    eng.init_device(seed)
    eng.set_trace(0)
    for t in range(102, 102 + steps):
        eng.run(t, t)
        w.modelrun(t, t, "probabilistic", 0)
        st = eng.download_compact(0)
        okH = np.array_equal(st["holdings"], w.holdings)
        okN = np.array_equal(st["nav_close"], np.asarray(w.nav_close), equal_nan=True)
        okP = np.array_equal(st["price_now"], w.stock_value[:, t - 1])
        if not (okH and okN and okP):
            ev = w.trace(); e = ev[ev["t"] == t]
            ch = e[e["kind"] == orc.EV_CHOICE]
            print(f"MISMATCH at t={t}: holdings {okH} nav_close {okN} prices {okP}; divest {int((e['kind']==orc.EV_DIVEST).sum())} "
                  f"spawn {int((e['kind']==orc.EV_SPAWN).sum())} reinvestors {len(ch)} distinct horizons {len(set(hz[ch['a']]))}")
            bad = np.nonzero(st["nav_close"] != np.asarray(w.nav_close))[0]
            print(" nav_close differs at funds", bad[:10], "n=", len(bad))
            print(" choices (oracle):", list(zip(ch["a"][:12], ch["b"][:12])))
            print(" engine status", eng.rep_status())
            tr = eng.download_trace(0)["events"]; te = tr[tr["t"] == t]
            print(" engine events kinds:", np.bincount(te["kind"], minlength=9), " oracle:", np.bincount(e["kind"], minlength=9))
            tch = te[te["kind"] == 7]
            print(" choices (engine):", list(zip(tch["a"][:12], tch["b"][:12])))
            print(" inj engine", tch["v"][:12], "\n inj oracle", ch["v"][:12])
            dH = st["holdings"] - w.holdings; rows = np.nonzero(np.abs(dH).sum(axis=1))[0]
            print(" holdings rows differing:", rows, "max rel", [float(np.max(np.abs(dH[r]) / np.maximum(np.abs(w.holdings[r]), 1e-300))) for r in rows[:8]])
            dP = st["price_now"] - w.stock_value[:, t - 1]; print(" prices differing:", int((dP != 0).sum()), "max rel", float(np.max(np.abs(dP) / w.stock_value[:, t - 1])))
            print(" NaN prices engine/oracle:", int(np.isnan(st['price_now']).sum()), int(np.isnan(w.stock_value[:, t-1]).sum()), " NaN nav_close engine/oracle:", int(np.isnan(st['nav_close']).sum()), int(np.isnan(np.asarray(w.nav_close)).sum()), " NaN nav_open engine", int(np.isnan(st['nav_open']).sum()))
            print(" nav_close engine at chosen funds:", st['nav_close'][ch['b']], " oracle:", np.asarray(w.nav_close)[ch['b']])
            print(" horizons of reinvestors:", list(hz[ch["a"]]))
            dv = e[e["kind"] == 2]; print(" divest funds (oracle)", list(dv["b"]))
            break
    else:
        print("all", steps, "steps equal")
