"""Generate tests/golden/*.npz by running the REFERENCE's own modules (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run as ``python -m oracle.make_golden``
from the repo root; needs /root/reference.  Every fixture records the outputs of
the reference modules imported by oracle/ref_loader.py (VecDGCNN_att, Shape_Prior,
FieldWrapper+DeepSDF_Decoder, sequential_matcher, nn_matcher,
kabsch_transformation_estimation) on seeded synthetic inputs, for
  * the shipped checkpoint   (suffix _shipped; needs the extracted weights on the GPU box)
  * seeded random weights    (suffix _random;  regenerated from the seed anywhere)
and cross-checks oracle/restatement.py against them before writing.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import ref_loader
from . import restatement as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
RANDOM_SEED = 0
FEAT_STRIDE = 16  # per-layer features are stored at every 16th point only


def _np(t):
    return t.detach().cpu().numpy()


def _relerr(a, b):
    return float((a - b).abs().max() / a.abs().max().clamp_min(1e-30))


def _hook_knn(mods, store):
    """Record the idx every knn_points / sample_farthest_points call of the reference returns."""
    enc_mod = mods.vec_dgcnn_atten
    real_knn, real_fps = enc_mod.knn_points, enc_mod.sample_farthest_points

    def knn(*a, **k):
        out = real_knn(*a, **k)
        store["knn_idx"].append(out[1].clone())
        return out

    def fps(*a, **k):
        out = real_fps(*a, **k)
        store["fps_idx"].append(out[1].clone())
        return out

    enc_mod.knn_points, enc_mod.sample_farthest_points = knn, fps
    return real_knn, real_fps


def encoder_fixture(sp, sd, tag, B, N, seed):
    mods = ref_loader.load()
    x = R.synth_instances(B, N, seed)
    store = {"knn_idx": [], "fps_idx": []}
    feats = []
    real = _hook_knn(mods, store)
    hooks = []
    # layer outputs == inputs of the next layer's V module / of conv_c
    try:
        with torch.no_grad():
            code = sp.encode(x)
            store_enc = {"knn_idx": list(store["knn_idx"]), "fps_idx": list(store["fps_idx"])}
            # direct encoder call on the normalised cloud (the VecDGCNN_att.forward contract)
            mu = x.mean(-1, keepdim=True)
            xc = x - mu
            s0 = R.scale0(xc)
            xn = xc / s0[:, None, None]
            store["knn_idx"].clear(), store["fps_idx"].clear()
            center, scale, z_so3, z_inv = sp.encoder(xn)
    finally:
        mods.vec_dgcnn_atten.knn_points, mods.vec_dgcnn_atten.sample_farthest_points = real
        for h in hooks:
            h.remove()
    # restatement cross-check (same graph expected; teacher-force nothing)
    tr = {}
    with torch.no_grad():
        c2, s2, zs2, zi2 = R.encoder_forward(sd, xn, trace=tr)
        code2 = R.encode(sd, x)
    for i, (a, b) in enumerate(zip(store["knn_idx"], tr["knn_idx"])):
        assert torch.equal(a, b), f"[{tag}] restatement kNN idx differs from reference at layer {i}"
    for i, (a, b) in enumerate(zip(store["fps_idx"], tr["fps_idx"])):
        assert torch.equal(a, b), f"[{tag}] restatement FPS idx differs at call {i}"
    errs = dict(center=_relerr(center, c2), scale=_relerr(scale, s2), z_so3=_relerr(z_so3, zs2),
                z_inv=_relerr(z_inv, zi2))
    errs.update({"enc_" + k: _relerr(code[k], code2[k]) for k in code})
    print(f"[{tag}] restatement vs reference max-rel: " + ", ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert max(errs.values()) < 2e-5, errs
    out = dict(
        x=_np(x), x_norm=_np(xn), scale0=_np(s0), seed=np.int64(seed),
        center=_np(center), scale=_np(scale), z_so3=_np(z_so3), z_inv=_np(z_inv),
        enc_z_so3=_np(code["z_so3"]), enc_z_inv=_np(code["z_inv"]), enc_s=_np(code["s"]), enc_t=_np(code["t"]),
    )
    for i, idx in enumerate(store["knn_idx"]):
        out[f"knn_idx_{i}"] = _np(idx).astype(np.int16)
    for i, idx in enumerate(store["fps_idx"]):
        out[f"fps_idx_{i}"] = _np(idx).astype(np.int16)
    for i, f in enumerate(tr["feat"]):  # restatement features (verified end-to-end above)
        out[f"feat_{i}"] = _np(f[..., ::FEAT_STRIDE].contiguous())
    # near-tie report: relative gap between the K-th and (K+1)-th neighbour in fp64
    gaps = []
    for i in range(len(tr["knn_idx"])):
        sf, df = tr["src_f"][i], tr["dst_f"][i]
        Bq, C, _, Nd = df.shape
        for b in range(Bq):
            d = R.p3d_shim._sqdist_f64(df[b].reshape(C * 3, Nd).T.contiguous(),
                                       sf[b].reshape(C * 3, -1).T.contiguous())
            dv, _ = torch.sort(d, dim=-1)
            if dv.shape[1] > 16:
                gaps.append(float(((dv[:, 16] - dv[:, 15]) / dv[:, 16].clamp_min(1e-300)).min()))
    out["min_rel_gap_k16"] = np.float64(min(gaps) if gaps else 1.0)
    np.savez_compressed(os.path.join(OUT, f"encoder_{tag}.npz"), **out)
    return x, code


def pair_fixture(sp, sd, tag, n_inst, N, seed):
    """C3-shaped: second set = permuted, rotated, translated, re-noised copy of the first."""
    mods = ref_loader.load()
    g = torch.Generator().manual_seed(seed + 77)
    xa = R.synth_instances(n_inst, N, seed)
    perm = torch.randperm(n_inst, generator=g)
    Rg = R.random_rotations(n_inst, seed + 1)
    tg = torch.randn(n_inst, 3, 1, generator=g)
    xb = Rg @ xa[perm] + tg + 0.002 * torch.randn(n_inst, 3, N, generator=g)
    with torch.no_grad():
        ca, cb = sp.encode(xa), sp.encode(xb)
        m = mods.matcher_new.sequential_matcher(ca["z_inv"], cb["z_inv"])
        mn = mods.matcher_new.nn_matcher(ca["z_inv"].T[None], cb["z_inv"].T[None])
        m0 = m["matches0"]
        # pose per matched pair exactly as more_solver.py:114-116 (z_so3 + t of each code)
        x1 = ca["z_so3"] + ca["t"]
        x2 = (cb["z_so3"] + cb["t"])[m0]
        Rk, tk, res, flag = mods.pose_estimation.kabsch_transformation_estimation(x1, x2)
        # restatement cross-check
        ca2, cb2 = R.encode(sd, xa), R.encode(sd, xb)
        m2 = R.sequential_match(ca2["z_inv"], cb2["z_inv"])
        assert torch.equal(m2["matches0"], m0) and torch.equal(m2["matches1"], m["matches1"])
        mn2 = R.mutual_nn_match(ca2["z_inv"].T[None], cb2["z_inv"].T[None])
        assert torch.equal(mn2["matches0"], mn["matches0"]) and torch.equal(mn2["matches1"], mn["matches1"])
        R2, t2, res2 = R.kabsch(ca2["z_so3"] + ca2["t"], (cb2["z_so3"] + cb2["t"])[m0])
    print(f"[{tag}] pair: matches0={m0.tolist()} (gt inverse-perm={torch.argsort(perm).tolist()}) "
          f"R err vs restatement {float((Rk - R2).abs().max()):.2e}, t {float((tk - t2).abs().max()):.2e}")
    # random weights give a nearly degenerate z_so3 cloud: the 3x3 problem is ill-conditioned and
    # amplifies the 5e-6 embedding difference; end-to-end pose parity is judged on the shipped weights.
    assert float((Rk - R2).abs().max()) < (5e-4 if tag.startswith("shipped") else 2e-2) and not flag
    np.savez_compressed(
        os.path.join(OUT, f"pair_{tag}.npz"),
        xa=_np(xa), xb=_np(xb), perm=_np(perm), R_gt=_np(Rg), t_gt=_np(tg),
        za_inv=_np(ca["z_inv"]), zb_inv=_np(cb["z_inv"]), za_so3=_np(ca["z_so3"]), zb_so3=_np(cb["z_so3"]),
        sa=_np(ca["s"]), sb=_np(cb["s"]), ta=_np(ca["t"]), tb=_np(cb["t"]),
        matches0=_np(m0), matches1=_np(m["matches1"]),
        nn_matches0=_np(mn["matches0"]), nn_matches1=_np(mn["matches1"]),
        R=_np(Rk), t=_np(tk), res=_np(res))


def sdf_fixture(sp, sd, tag, x, code, M, seed):
    g = torch.Generator().manual_seed(seed)
    B = x.shape[0]
    # queries uniform in the 1.1-padded unit cube of each instance's canonical frame
    # (mesh_extractor2.py:100), mapped to world by q*s + t; plus the input points themselves.
    qc = (torch.rand(B, M, 3, generator=g) - 0.5) * 1.1
    q = qc * code["s"][:, None, None] + code["t"]
    q = torch.cat([q, x.transpose(1, 2)[:, :256]], 1)
    with torch.no_grad():
        sdf = sp.decoder(q, None, code, return_sdf=True)
        occ_logits = sp.decoder(q, None, code).logits
        sdf2 = R.sdf_decode(sd, q, code)
    err = float((sdf - sdf2).abs().max())
    print(f"[{tag}] sdf: restatement max-abs err {err:.2e}; mean |sdf| {float(sdf.abs().mean()):.3f}")
    assert err < 5e-6 and torch.equal(occ_logits, -sdf)
    np.savez_compressed(os.path.join(OUT, f"sdf_{tag}.npz"), query=_np(q), sdf=_np(sdf),
                        z_so3=_np(code["z_so3"]), z_inv=_np(code["z_inv"]), s=_np(code["s"]), t=_np(code["t"]))


def solver_fixture():
    """Weight-free cases for the matchers and Kabsch, including the reference's quirks."""
    mods = ref_loader.load()
    g = torch.Generator().manual_seed(4242)
    out = {}
    cases = [(7, 9), (9, 7), (32, 32), (1, 5), (5, 1), (16, 16), (40, 23)]
    for ci, (n, m) in enumerate(cases):
        z0 = torch.randn(n, 256, generator=g)
        z1 = torch.randn(m, 256, generator=g)
        if ci == 2:  # permuted noisy copies: the realistic case
            z1 = z0[torch.randperm(n, generator=g)] + 0.05 * torch.randn(n, 256, generator=g)
        if ci == 5:  # all-negative scores: exercises the "/(max+1e-5)" sign-flip quirk (matcher_new.py:123)
            base = torch.randn(256, generator=g)
            z0 = base[None] + 0.3 * torch.randn(n, 256, generator=g)
            z1 = -base[None] + 0.3 * torch.randn(m, 256, generator=g)
        if ci == 6:  # exact duplicates: tie-breaking by row-major first hit
            z1[3] = z1[11]
            z0[5] = z0[17]
        r = mods.matcher_new.sequential_matcher(z0, z1)
        rn = mods.matcher_new.nn_matcher(z0.T[None], z1.T[None])
        r2 = R.sequential_match(z0, z1)
        rn2 = R.mutual_nn_match(z0.T[None], z1.T[None])
        for k in ("matches0", "matches1"):
            assert torch.equal(r[k], r2[k]), (ci, k)
            assert torch.equal(rn[k].reshape(-1), rn2[k].reshape(-1)), (ci, k)
        out[f"m{ci}_z0"], out[f"m{ci}_z1"] = _np(z0), _np(z1)
        out[f"m{ci}_seq0"], out[f"m{ci}_seq1"] = _np(r["matches0"]), _np(r["matches1"])
        out[f"m{ci}_nn0"], out[f"m{ci}_nn1"] = _np(rn["matches0"].reshape(-1)), _np(rn["matches1"].reshape(-1))
    out["n_match_cases"] = np.int64(len(cases))
    # Kabsch: generic, noisy, reflected (det fix), planar, weighted
    b, n = 6, 256
    x1 = torch.randn(b, n, 3, generator=g)
    Rg = R.random_rotations(b, 99)
    tg = torch.randn(b, 1, 3, generator=g)
    x2 = x1 @ Rg.transpose(1, 2) + tg
    x2[1] += 0.05 * torch.randn(n, 3, generator=g)
    x2[2] = x2[2] * torch.tensor([1.0, 1.0, -1.0])      # mirrored target: forces det = -1 branch
    x1[3, :, 2] = 0.0                                      # planar source
    x2[3] = x1[3] @ Rg[3].T + tg[3]
    x1[4] *= 1e-3
    x2[4] = x1[4] @ Rg[4].T + tg[4]                        # tiny scale
    w = torch.rand(b, n, generator=g)
    Rk, tk, res, flag = mods.pose_estimation.kabsch_transformation_estimation(x1, x2)
    Rw, tw, resw, _ = mods.pose_estimation.kabsch_transformation_estimation(x1, x2, weights=w.clone())
    R2, t2, res2 = R.kabsch(x1, x2)
    Rw2, tw2, resw2 = R.kabsch(x1, x2, weights=w.clone())
    print(f"[solver] kabsch restatement err R {float((Rk - R2).abs().max()):.2e} "
          f"t {float((tk - t2).abs().max()):.2e} weighted R {float((Rw - Rw2).abs().max()):.2e}")
    out.update(k_x1=_np(x1), k_x2=_np(x2), k_w=_np(w), k_R=_np(Rk), k_t=_np(tk), k_res=_np(res),
               k_Rw=_np(Rw), k_tw=_np(tw), k_resw=_np(resw), k_R_gt=_np(Rg))
    np.savez_compressed(os.path.join(OUT, "solver_cases.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    assert ref_loader.available() and ref_loader.checkpoint_available(), "needs /root/reference"
    sd_ship = ref_loader.shipped_state_dict()
    sd_rand = R.random_state_dict(RANDOM_SEED)
    only = sys.argv[1:]
    if not only or "solver" in only:
        solver_fixture()
    for tag, sd in (("shipped", sd_ship), ("random", sd_rand)):
        if only and tag not in only:
            continue
        sp = ref_loader.shape_prior(None if tag == "shipped" else sd)
        x, code = encoder_fixture(sp, sd, tag, B=2, N=1024, seed=1234)
        sdf_fixture(sp, sd, tag, x, code, M=1792, seed=1239)
        encoder_fixture(sp, sd, tag + "_n2048", B=1, N=2048, seed=1236)
        pair_fixture(sp, sd, tag, n_inst=4, N=2048, seed=1236)
    print("fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KB")


if __name__ == "__main__":
    sys.exit(main())
