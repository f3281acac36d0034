"""Config-scale golden fixtures (BASELINE.json configs C2 / C3 / C5) from the REFERENCE's own modules.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``python -m oracle.make_golden_configs [c2] [c3] [c5]`` from the
repo root, build container only (needs /root/reference and its shipped checkpoint; ~15 min of CPU).  Inputs are
the asymmetric ``synth_parts`` family (livingscenes_b200/synthetic.py); everything stored was produced by the
reference's Shape_Prior / sequential_matcher / nn_matcher / kabsch_transformation_estimation / FieldWrapper.

  c2_encoder_shipped.npz  16 instances x 1024 points: encode dict, all 7 kNN graphs, the 3 FPS selections
  c3_pair_shipped.npz     32 + 32 instances x 2048 points (permuted, rotated, translated, re-noised rescan):
                          codes of both scans, sequential + mutual-NN matches, Kabsch R / t / residuals
  c5_sdf_shipped.npz      4 codes x 100 000 queries (regenerated from the seed by synthetic.sdf_queries): SDF,
                          and the reference's own |sdf| < 0.01 shell sizes
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from livingscenes_b200 import synthetic as S

from . import ref_loader
from . import restatement as R
from .make_golden import OUT, _hook_knn, _np

C2_SEED, C3_SEED, C5_SEED = 1235, 1236, 1239


def _encode_chunks(sp, x, chunk=8):
    outs = [sp.encode(x[i:i + chunk]) for i in range(0, x.shape[0], chunk)]
    return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}


def c2(sp, sd):
    mods = ref_loader.load()
    x = S.synth_parts(16, 1024, C2_SEED)
    store = {"knn_idx": [], "fps_idx": []}
    real = _hook_knn(mods, store)
    try:
        with torch.no_grad():
            code = sp.encode(x)
    finally:
        mods.vec_dgcnn_atten.knn_points, mods.vec_dgcnn_atten.sample_farthest_points = real
    # restatement cross-check.  A single fp32 near-tie in a feature-space kNN row flips a neighbour and moves the
    # embedding by ~1e-3 (the network amplifies it), so instances are compared per graph: where the restatement
    # built the reference's graphs the codes must agree to 2e-5; elsewhere it is re-run on the reference's graphs.
    tr = {}
    with torch.no_grad():
        code2 = R.encode(sd, x, trace=tr)
    same = torch.ones(x.shape[0], dtype=torch.bool)
    for a, b in zip(store["knn_idx"], tr["knn_idx"]):
        same &= (a.sort(-1)[0] == b.sort(-1)[0]).all(-1).all(-1)
    err = torch.stack([(code[k] - code2[k]).reshape(x.shape[0], -1).abs().amax(1) /
                       code[k].reshape(x.shape[0], -1).abs().amax(1) for k in code]).amax(0)
    print(f"[c2] restatement: graphs identical for {int(same.sum())}/{x.shape[0]} instances; max-rel on those "
          f"{float(err[same].max()):.2e}; on the near-tie-flipped ones {float(err[~same].max()) if (~same).any() else 0:.2e}")
    assert float(err[same].max()) < 2e-5
    if (~same).any():
        with torch.no_grad():
            c_, s_, zs_, zi_ = R.encoder_forward(sd, tr["x_norm"][~same], force={"knn_idx": [t[~same] for t in store["knn_idx"]],
                                                                 "fps_idx": [t[~same] for t in store["fps_idx"]]})
        e2 = float((zs_ - code["z_so3"][~same]).abs().max() / code["z_so3"].abs().max())
        print(f"[c2] restatement driven with the reference graphs on the flipped instances: z_so3 max-rel {e2:.2e}")
        assert e2 < 2e-5
    out = dict(x=_np(x), z_so3=_np(code["z_so3"]), z_inv=_np(code["z_inv"]), s=_np(code["s"]), t=_np(code["t"]))
    for i, idx in enumerate(store["knn_idx"]):
        out[f"knn_idx_{i}"] = _np(idx).astype(np.int16)
    for i, idx in enumerate(store["fps_idx"]):
        out[f"fps_idx_{i}"] = _np(idx).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "c2_encoder_shipped.npz"), **out)
    return x, code


def c3(sp, sd):
    mods = ref_loader.load()
    n, N = 32, 2048
    g = torch.Generator().manual_seed(C3_SEED + 77)
    xa = S.synth_parts(n, N, C3_SEED)
    perm = torch.randperm(n, generator=g)
    Rg = S.random_rotations(n, C3_SEED + 1)
    tg = torch.randn(n, 3, 1, generator=g)
    xb = Rg @ xa[perm] + tg + 0.002 * torch.randn(n, 3, N, generator=g)
    with torch.no_grad():
        ca, cb = _encode_chunks(sp, xa, 4), _encode_chunks(sp, xb, 4)
        m = mods.matcher_new.sequential_matcher(ca["z_inv"], cb["z_inv"])
        mn = mods.matcher_new.nn_matcher(ca["z_inv"].T[None], cb["z_inv"].T[None])
        m0 = m["matches0"]
        Rk, tk, res, flag = mods.pose_estimation.kabsch_transformation_estimation(
            ca["z_so3"] + ca["t"], (cb["z_so3"] + cb["t"])[m0])
    inv = torch.argsort(perm)
    cos = ((torch.einsum("bij,bij->b", Rk, Rg[inv]) - 1) / 2).clamp(-1, 1)
    print(f"[c3] recall vs planted permutation {float((m0 == inv).float().mean()):.3f}; "
          f"median RRE {float(torch.rad2deg(torch.acos(cos)).median()):.2f} deg; flag {flag}")
    np.savez_compressed(
        os.path.join(OUT, "c3_pair_shipped.npz"),
        xa=_np(xa), xb=_np(xb), perm=_np(perm), R_gt=_np(Rg), t_gt=_np(tg),
        za_inv=_np(ca["z_inv"]), zb_inv=_np(cb["z_inv"]), za_so3=_np(ca["z_so3"]), zb_so3=_np(cb["z_so3"]),
        sa=_np(ca["s"]), sb=_np(cb["s"]), ta=_np(ca["t"]), tb=_np(cb["t"]),
        matches0=_np(m0), matches1=_np(m["matches1"]),
        nn_matches0=_np(mn["matches0"].reshape(-1)), nn_matches1=_np(mn["matches1"].reshape(-1)),
        R=_np(Rk), t=_np(tk), res=_np(res))


def c5(sp, sd, x, code):
    B, M = 4, 100_000
    code = {k: v[:B] for k, v in code.items()}
    q = S.sdf_queries(code["s"], code["t"], M, C5_SEED)
    sdf = []
    with torch.no_grad():
        for m0 in range(0, M, 10000):  # the reference's points_batch_size (mesh_extractor2.py:133-156)
            sdf.append(sp.decoder(q[:, m0:m0 + 10000], None, code, return_sdf=True))
        sdf = torch.cat(sdf, 1)
        sdf2 = R.sdf_decode(sd, q[:, :4096], code)
    err = float((sdf[:, :4096] - sdf2).abs().max())
    shell = (sdf.abs() < 0.01).sum(1)
    print(f"[c5] restatement max-abs err {err:.2e}; shell sizes {shell.tolist()}; mean |sdf| {float(sdf.abs().mean()):.3f}")
    assert err < 5e-6
    np.savez_compressed(os.path.join(OUT, "c5_sdf_shipped.npz"), sdf=_np(sdf), seed=np.int64(C5_SEED),
                        M=np.int64(M), z_so3=_np(code["z_so3"]), z_inv=_np(code["z_inv"]), s=_np(code["s"]),
                        t=_np(code["t"]), shell=_np(shell))


def matchers2():
    """Secondary matchers (matcher_new.py:45-71,142-230) run by the reference's own functions on the C3 codes and on
    seeded random cases.  ``sinkhorn_matcher`` hard-codes ``.cuda()`` for its alpha scalar: patched to a no-op here."""
    from unittest import mock

    mods = ref_loader.load()
    g3 = dict(np.load(os.path.join(OUT, "c3_pair_shipped.npz")))
    gen = torch.Generator().manual_seed(99)
    cases = [({"z_inv": torch.from_numpy(g3["za_inv"]), "z_so3": torch.from_numpy(g3["za_so3"])},
              {"z_inv": torch.from_numpy(g3["zb_inv"]), "z_so3": torch.from_numpy(g3["zb_so3"])})]
    for n, m in ((7, 9), (12, 5), (20, 20)):
        za = {"z_inv": torch.randn(n, 256, generator=gen), "z_so3": torch.randn(n, 256, 3, generator=gen)}
        perm = torch.randperm(max(n, m), generator=gen)[:m] % n
        Rr = S.random_rotations(m, 5 + n)
        zb = {"z_inv": za["z_inv"][perm] + 0.3 * torch.randn(m, 256, generator=gen),
              "z_so3": za["z_so3"][perm] @ Rr.transpose(1, 2) + 0.05 * torch.randn(m, 256, 3, generator=gen)}
        cases.append((za, zb))
    out = {"n_cases": np.int64(len(cases))}
    with mock.patch.object(torch.Tensor, "cuda", lambda self, *a, **k: self):
        for ci, (za, zb) in enumerate(cases):
            sk = mods.matcher_new.sinkhorn_matcher(za["z_inv"].T[None], zb["z_inv"].T[None])
            s3 = mods.matcher_new.sim3_seq_matcher(za, zb)
            eq = mods.matcher_new.eq_seq_matcher(za, zb)
            sk2 = R.sinkhorn_match(za["z_inv"].T[None], zb["z_inv"].T[None])
            s32, eq2 = R.residual_seq_match(za, zb, True), R.residual_seq_match(za, zb, False)
            for k in ("matches0", "matches1"):
                assert torch.equal(sk[k].reshape(-1), sk2[k].reshape(-1)), ("sinkhorn", ci, k)
                assert torch.equal(s3[k], s32[k]), ("sim3_seq", ci, k)
                assert torch.equal(eq[k], eq2[k]), ("eq_seq", ci, k)
            for nm, t in (("a_inv", za["z_inv"]), ("a_so3", za["z_so3"]), ("b_inv", zb["z_inv"]), ("b_so3", zb["z_so3"])):
                out[f"c{ci}_{nm}"] = _np(t)
            for nm, r in (("sk", sk), ("s3", s3), ("eq", eq)):
                out[f"c{ci}_{nm}0"], out[f"c{ci}_{nm}1"] = _np(r["matches0"].reshape(-1)), _np(r["matches1"].reshape(-1))
            print(f"[matchers2] case {ci}: sinkhorn matched {int((sk['matches0'] >= 0).sum())}, sim3_seq/eq_seq agree with "
                  f"sequential on {int((s3['matches0'] == eq['matches0']).sum())}/{len(s3['matches0'])}")
    np.savez_compressed(os.path.join(OUT, "matchers2_cases.npz"), **out)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    assert ref_loader.available() and ref_loader.checkpoint_available(), "needs /root/reference"
    only = sys.argv[1:]
    sd = ref_loader.shipped_state_dict()
    sp = ref_loader.shape_prior(None) if (not only or set(only) & {"c2", "c3", "c5"}) else None
    t0 = time.time()
    x = code = None
    if not only or "c2" in only or "c5" in only:
        x, code = c2(sp, sd)
        print(f"c2 done {time.time() - t0:.0f}s")
    if not only or "c5" in only:
        c5(sp, sd, x, code)
        print(f"c5 done {time.time() - t0:.0f}s")
    if not only or "c3" in only:
        c3(sp, sd)
        print(f"c3 done {time.time() - t0:.0f}s")
    if not only or "matchers2" in only:
        matchers2()
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.0f} KB")


if __name__ == "__main__":
    sys.exit(main())
