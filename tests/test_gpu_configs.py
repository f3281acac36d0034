"""Config-scale parity (BASELINE.json configs C2 / C3 / C5) against fixtures produced by the REFERENCE's own
modules (oracle/make_golden_configs.py) and the planted-transform sanity check of the synthetic workload.
All calls go through the C ABI (ctypes) -- see tests/test_gpu_parity.py for the kernel-level tests."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden, relerr, state_dict_for

pytestmark = pytest.mark.gpu
TOL = 1e-4  # BASELINE.json north_star: poses and SDF within 1e-4 relative fp32


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def shipped(dev):
    import livingscenes_b200 as ls

    return ls.Shape_Prior.from_state_dict(state_dict_for("shipped")).to(dev).eval()


def _need(name):
    if not os.path.exists(os.path.join(GOLDEN, name + ".npz")):
        pytest.skip(f"{name}.npz not generated")
    return golden(name)


def _codes_match_modulo_near_ties(model, x, ref, min_exact_frac=0.7):
    """Free-running parity at config scale.  One fp32 near-tie in a feature-space kNN row flips a neighbour and moves
    the embedding by ~1e-3 (measured between the reference and our CPU restatement of it, oracle/make_golden_configs.py),
    so instances split in two groups: codes within 1e-4 of the reference's (required for >= min_exact_frac of them),
    and the rest, which must (a) equal the ORACLE driven with the graph the CUDA path actually built (all the
    arithmetic, given the graph) and (b) differ from the oracle's own graph only in fp32 near-tie rows.
    Returns the boolean mask of the first group."""
    from oracle import restatement as R
    from test_gpu_parity import _knn_rows_equivalent

    r = model.encoder.run(x, normalize=True, taps=True)
    code = {"z_so3": r["z_so3"], "z_inv": r["z_inv"], "s": r["scale"], "t": r["center"].unsqueeze(1)}
    torch.cuda.synchronize()
    B = x.shape[0]
    err = torch.stack([(code[k].cpu() - torch.as_tensor(ref[k])).reshape(B, -1).abs().amax(1) /
                       torch.as_tensor(ref[k]).reshape(B, -1).abs().amax(1).clamp_min(1e-30) for k in code]).amax(0)
    ok = err < TOL
    assert float(ok.float().mean()) >= min_exact_frac, f"only {int(ok.sum())}/{B} instances within 1e-4: {err.tolist()}"
    bad = (~ok).nonzero().reshape(-1)
    if bad.numel():
        sd = state_dict_for("shipped")
        xn = r["x_norm"][bad].cpu()
        tr = {}
        with torch.no_grad():
            c, sc, zs, zi = R.encoder_forward(sd, xn, force={"knn_idx": [t[bad].cpu() for t in r["knn_idx"]],
                                                             "fps_idx": [t[bad].cpu() for t in r["fps_idx"]]})
            R.encoder_forward(sd, xn, trace=tr)  # the oracle's own graph on the same normalised clouds
        # run(normalize=True) returns s = scale_0 * scale (Shape_Prior.encode); the oracle call above is the bare encoder
        assert relerr(r["z_so3"][bad], zs) < TOL and relerr(r["z_inv"][bad], zi) < TOL
        assert relerr(r["scale"][bad].cpu() / r["scale0"][bad].cpu(), sc) < TOL
        # the first layer whose graph differs must differ in near-tie rows only (inputs still agree to ~1e-6 there)
        for j, b in enumerate(bad.tolist()):
            for i in range(7):
                mine, theirs = r["knn_idx"][i][b:b + 1].cpu(), tr["knn_idx"][i][j:j + 1]
                if (mine.sort(-1)[0] != theirs.sort(-1)[0]).any():
                    nbad, ndiff, _ = _knn_rows_equivalent(mine, theirs, tr["dst_f"][i][j:j + 1], tr["src_f"][i][j:j + 1],
                                                          rel_tol=3e-5)
                    assert nbad == 0, f"instance {b} layer {i}: {nbad} of {ndiff} differing rows are not near-ties"
                    break
    return ok, r, code


def test_c2_encoder_16_instances_matches_reference(shipped, dev):
    """C2 scale: 16 shipped-weight instances of 1024 points in ONE batch against the reference's own run: FPS
    selections exact, kNN graphs exact for every instance that is not near-tie-flipped, codes within 1e-4."""
    g = _need("c2_encoder_shipped")
    x = torch.from_numpy(g["x"]).to(dev)
    ref = {k: g[k] for k in ("z_so3", "z_inv", "s", "t")}
    ok, r, _ = _codes_match_modulo_near_ties(shipped, x, ref)
    assert len(r["fps_idx"]) == 3
    for i, sel in enumerate(r["fps_idx"]):
        assert np.array_equal(sel.cpu().numpy(), g[f"fps_idx_{i}"].astype(np.int64)), f"FPS level {i}"
    n_same = 0
    for b in range(x.shape[0]):
        same = all(np.array_equal(np.sort(r["knn_idx"][i][b].cpu().numpy(), -1),
                                  np.sort(g[f"knn_idx_{i}"][b].astype(np.int64), -1)) for i in range(7))
        n_same += same
    print(f"[c2] instances with all 7 kNN graphs identical to the reference run: {n_same}/16; codes within 1e-4: {int(ok.sum())}/16")
    assert n_same >= 11


def test_c3_pair_32x32_n2048_matches_reference(shipped, dev):
    """BASELINE config 3 at full size: 2 x 32 instances x 2048 points, shipped weights, against the reference's own
    run: bit-exact sequential and mutual-NN assignments, codes / R / t within 1e-4 (modulo near-tie graph flips)."""
    import livingscenes_b200 as ls

    g = _need("c3_pair_shipped")
    xa, xb = torch.from_numpy(g["xa"]).to(dev), torch.from_numpy(g["xb"]).to(dev)
    oka, _, ca = _codes_match_modulo_near_ties(shipped, xa, {"z_so3": g["za_so3"], "z_inv": g["za_inv"], "s": g["sa"], "t": g["ta"]})
    okb, _, cb = _codes_match_modulo_near_ties(shipped, xb, {"z_so3": g["zb_so3"], "z_inv": g["zb_inv"], "s": g["sb"], "t": g["tb"]})
    print(f"[c3] instances within 1e-4 of the reference: ref scan {int(oka.sum())}/32, rescan {int(okb.sum())}/32")
    solver = ls.More_Solver(shipped)
    # kernel level, on the reference's own embeddings: assignments bit-exact, poses 1e-4
    gca = {"z_so3": torch.from_numpy(g["za_so3"]).to(dev), "z_inv": torch.from_numpy(g["za_inv"]).to(dev),
           "t": torch.from_numpy(g["ta"]).to(dev)}
    gcb = {"z_so3": torch.from_numpy(g["zb_so3"]).to(dev), "z_inv": torch.from_numpy(g["zb_inv"]).to(dev),
           "t": torch.from_numpy(g["tb"]).to(dev)}
    m = solver._solve_object_matching(gca, gcb, "sequential")
    assert np.array_equal(m["matches0"].cpu().numpy(), g["matches0"]) and np.array_equal(m["matches1"].cpu().numpy(), g["matches1"])
    nn = solver._solve_object_matching(gca, gcb, "nn")
    assert np.array_equal(nn["matches0"].reshape(-1).cpu().numpy(), g["nn_matches0"].reshape(-1))
    assert np.array_equal(nn["matches1"].reshape(-1).cpu().numpy(), g["nn_matches1"].reshape(-1))
    R, t, res = ls.kabsch_from_codes(gca, gcb, torch.from_numpy(g["matches0"]).to(dev))
    assert float((R.cpu() - torch.from_numpy(g["R"])).abs().max()) < TOL
    assert relerr(t, g["t"]) < TOL and relerr(res, g["res"]) < 1e-3
    # end to end (own embeddings): same assignments; poses of the pairs whose two instances are un-flipped
    out = solver.solve_scene_pair(xa, xb)
    torch.cuda.synchronize()
    assert np.array_equal(out["matches"]["matches0"].cpu().numpy(), g["matches0"])
    assert np.array_equal(out["matches"]["matches1"].cpu().numpy(), g["matches1"])
    pair_ok = oka & okb[torch.from_numpy(g["matches0"])]
    assert int(pair_ok.sum()) >= 16
    assert float((out["R"].cpu() - torch.from_numpy(g["R"]))[pair_ok].abs().max()) < 1e-3
    assert float((out["t"].cpu() - torch.from_numpy(g["t"]))[pair_ok].abs().max()) < 1e-3 * max(1.0, float(np.abs(g["t"]).max()))


def _chamfer(a, b):
    """symmetric mean squared nearest-neighbour distance (the form of evaluate.py:111-123 without transforms)."""
    d = torch.cdist(a.double(), b.double()) ** 2
    return float(d.min(1)[0].mean() + d.min(0)[0].mean())


def test_c5_sdf_4x100k_matches_reference(shipped, dev):
    """BASELINE config 5 scale per instance: 4 codes x 100 000 queries: max-abs SDF error <= 1e-4 and the
    |sdf| < 0.01 shells coincide with the reference's (Chamfer between shells ~ 0)."""
    from livingscenes_b200 import synthetic as S

    g = _need("c5_sdf_shipped")
    code = {k: torch.from_numpy(g[k]).to(dev) for k in ("z_so3", "z_inv", "s", "t")}
    q = S.sdf_queries(torch.from_numpy(g["s"]), torch.from_numpy(g["t"]), int(g["M"]), int(g["seed"]))
    sdf = shipped.decoder(q.to(dev), None, code, return_sdf=True).cpu()
    ref = torch.from_numpy(g["sdf"])
    assert float((sdf - ref).abs().max()) < TOL
    for b in range(q.shape[0]):
        mine, theirs = q[b][sdf[b].abs() < 0.01], q[b][ref[b].abs() < 0.01]
        assert abs(len(mine) - len(theirs)) <= max(2, len(theirs) // 200)
        if len(theirs) > 8:
            n = min(len(theirs), 4000)
            cd = _chamfer(mine[:n], theirs[:n])
            ext = float(g["s"][b]) ** 2
            assert cd < 1e-6 * ext + 1e-12 or cd / ext < 1e-4, (b, cd)


@pytest.mark.parametrize("n_inst,N", [(32, 2048), (32, 1024)])
def test_planted_transform_is_recovered(shipped, dev, n_inst, N):
    """Task sanity of the synthetic workload (bench.py / scripts/run_configs.py): the rescan is an exact rigid,
    permuted copy of asymmetric ``synth_parts`` objects; the path must recover the permutation and the poses."""
    import livingscenes_b200 as ls
    from livingscenes_b200 import synthetic as S

    xa = S.synth_parts(n_inst, N, 4242 + N)
    g = torch.Generator().manual_seed(7)
    perm = torch.randperm(n_inst, generator=g)
    Rg = S.random_rotations(n_inst, 4243)
    tg = torch.randn(n_inst, 3, 1, generator=g)
    xb = Rg @ xa[perm] + tg
    out = ls.More_Solver(shipped).solve_scene_pair(xa.to(dev), xb.to(dev))
    inv = torch.argsort(perm)
    m0 = out["matches"]["matches0"].cpu()
    assert torch.equal(m0, inv), f"match recall {float((m0 == inv).float().mean()):.3f}"
    R = out["R"].cpu()
    cos = ((torch.einsum("bij,bij->b", R, Rg[inv]) - 1) / 2).clamp(-1, 1)
    rre = torch.rad2deg(torch.acos(cos))
    rte = (out["t"].cpu() - tg[inv]).norm(dim=1).squeeze()
    assert float(rre.median()) < 1.0 and float(rre.max()) < 10.0, rre.tolist()
    assert float(rte.median()) < 0.05, rte.tolist()


def test_shape_prior_constructor_loads_reference_checkpoint(dev):
    """The real ``Shape_Prior(cfg, model_id)`` constructor (yaml + checkpoint, model_utils.py:83-163) against
    ``from_state_dict`` -- only where the reference tree is mounted (the build container)."""
    import livingscenes_b200 as ls
    from livingscenes_b200 import synthetic as S

    root = "/root/reference"
    if not os.path.exists(os.path.join(root, "weights/checkpoint/LivingScenes_latest.pt")):
        pytest.skip("reference tree not mounted")
    sp = ls.Shape_Prior({"working_dir": root, "field_cfg": "weights/files_backup/model_config.yaml",
                         "field_pt": "weights/checkpoint/LivingScenes_latest.pt"}, "chair", use_double=False)
    sp = sp.to(dev).eval()
    ref = ls.Shape_Prior.from_state_dict(state_dict_for("shipped")).to(dev).eval()
    x = S.synth_parts(2, 1024, 3).to(dev)
    a, b = sp.encode(x), ref.encode(x)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert sp.field_input_n == 1024 and sp.model_id == "chair"


@pytest.mark.parametrize("n_scene", [128, 1024])
def test_c4_sharded_scene_nccl(n_scene):
    """BASELINE config[3] on >= 2 GPUs over NCCL (skipped on a 1-GPU box; run with ``gpurun --gpus 2``): the gathered
    [n,1028] table and matches0 are bit-identical to a 1-GPU run of the full list (tests/dist_c4_worker.py)."""
    import subprocess
    import sys

    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n_gpu < 4 else 4
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29611 + (n_scene % 7)),
           os.path.join(here, "dist_c4_worker.py"), str(n_scene)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"C4 OK world={world} n={n_scene}" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
