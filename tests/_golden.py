"""Shared helpers: load a golden case (made by oracle/make_golden.py from the unmodified reference)
and replay it through any engine class with the reference's constructor / per-block call."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    cfg = {k[4:]: g[k].item() for k in g if k.startswith("cfg_")}
    ctor = {k[5:]: g[k].item() for k in g if k.startswith("ctor_")}
    return g, cfg, ctor


def rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def replay(engine_cls, name, extra_ctor=None, on_block=None):
    """Run the golden inputs through engine_cls; returns (engine, per-block dict of comparisons).

    Comparisons are relative L2 errors against what the reference produced."""
    g, cfg, ctor = load_case(name)
    kw = dict(ctor)
    kw.update(extra_ctor or {})
    np.random.seed(int(g["seed"]))
    eng = engine_cls(rir_A=g["rir_A"], rir_B=g["rir_B"], **cfg, **kw)
    nblk = int(g["nblk"])
    w_ranks = g["w_ranks"]
    out_ranks = g["out_ranks"]
    res = {}
    for t in range(nblk):
        outs = eng.process_input_buffers(g["input_A"][t], g["input_B"][t])
        if f"r_A_{t}" not in g:
            continue
        e = {}
        for z in ("A", "B"):
            if f"w_{z}_{t}" in g:
                w = getattr(eng, f"w_{z}")
                gw = g[f"w_{z}_{t}"]
                e[f"w_{z}"] = max(rel(w[v, :, 0], gw[i]) for i, v in enumerate(w_ranks))
                lam = np.asarray(getattr(eng, f"lambda_{z}"))[: len(g[f"lambda_{z}_{t}"])]
                e[f"lambda_{z}"] = float(np.max(np.abs(lam - g[f"lambda_{z}_{t}"])) / np.max(np.abs(g[f"lambda_{z}_{t}"])))
                e[f"r_{z}"] = rel(getattr(eng, f"r_{z}"), g[f"r_{z}_{t}"])
        for nm in ("R_A_to_A", "R_A_to_B", "R_B_to_A", "R_B_to_B"):
            if f"{nm}_diag_{t}" in g:
                R = np.asarray(getattr(eng, nm))
                n = R.shape[0]
                e[nm] = max(rel(np.diag(R), g[f"{nm}_diag_{t}"]), rel(R[[0, n // 2 - 1, n - 1], :], g[f"{nm}_rows_{t}"]))
        for i, nm in enumerate(("out_A", "out_B", "out_A_t", "out_B_t")):
            if f"{nm}_{t}" in g:
                o = outs[i]
                scale = max(np.linalg.norm(g[f"{nm}_{t}"]), 1e-300)
                e[nm] = float(max(np.linalg.norm(np.asarray(o[v]) - g[f"{nm}_{t}"][j]) for j, v in enumerate(out_ranks)) / scale)
        res[t] = e
        if on_block is not None:
            on_block(t, eng, outs)
    return eng, g, res


def compare_state(eng, g):
    out = {}
    for k in g:
        if k.startswith("state_"):
            a = k[6:]
            v = getattr(eng, a, None)
            if v is None:
                continue
            out[a] = rel(np.asarray(v), g[k]) if np.linalg.norm(g[k]) > 0 else float(np.linalg.norm(np.asarray(v)))
    return out
