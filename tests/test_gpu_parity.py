"""GPU parity tests proper: every call goes through the C ABI (ctypes -> _ls_b200.so) on cuda:0 and is
compared with the oracle (oracle/restatement.py) on the same seeded inputs and with the committed
golden fixtures generated from the reference's own modules.

Bars (BASELINE.json north_star): kNN / FPS indices and match assignments bit-exact (kNN up to
fp32 near-ties, which are counted and bounded); embeddings, poses, SDF within 1e-4 (relative to the
tensor's max-abs for embeddings/poses, absolute on the tanh output for SDF)."""
import numpy as np
import pytest
import torch

from conftest import golden, relerr, state_dict_for

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _model(tag, dev):
    import livingscenes_b200 as ls

    return ls.Shape_Prior.from_state_dict(state_dict_for(tag)).to(dev).eval()


def _knn_rows_equivalent(idx_cuda, idx_ref, dst_f, src_f, rel_tol=2e-6):
    """Rows where the CUDA neighbour list differs from the oracle's must be fp32 near-ties: the sorted
    fp64 distances of both selections agree to rel_tol.  Returns (n_bad_rows, n_diff_rows, n_rows)."""
    from oracle.p3d_shim import _sqdist_f64

    idx_cuda, idx_ref = idx_cuda.cpu(), idx_ref.cpu()
    B, Nd, K = idx_ref.shape
    diff_rows = (idx_cuda != idx_ref).any(-1)
    bad = 0
    for b in range(B):
        rows = diff_rows[b].nonzero().reshape(-1)
        if rows.numel() == 0:
            continue
        C = dst_f.shape[1]
        q = dst_f[b].reshape(C * 3, Nd).T[rows].contiguous()
        s = src_f[b].reshape(C * 3, -1).T.contiguous()
        d = _sqdist_f64(q, s)
        dc = torch.gather(d, 1, idx_cuda[b, rows]).sort(dim=1)[0]
        dr = torch.gather(d, 1, idx_ref[b, rows]).sort(dim=1)[0]
        ok = ((dc - dr).abs() <= rel_tol * dr.abs().clamp_min(1e-30)).all(-1)
        # the CUDA list must itself be ascending (up to the same tolerance) with distinct entries
        dcu = torch.gather(d, 1, idx_cuda[b, rows])
        asc = (dcu[:, 1:] >= dcu[:, :-1] * (1 - rel_tol) - 1e-30).all(-1)
        distinct = torch.tensor([len(set(r.tolist())) == K for r in idx_cuda[b, rows]])
        bad += int((~(ok & asc & distinct)).sum())
    return bad, int(diff_rows.sum()), B * Nd


# ------------------------------------------------------------------------------------------ graph ops
@pytest.mark.parametrize("tag", ["random", "shipped"])
def test_knn_kernel_teacher_forced(tag, dev, oracle_R):
    """ls_knn fed the oracle's layer inputs reproduces the oracle's graph at every layer."""
    from livingscenes_b200.ops import knn_graph_cm

    sd = state_dict_for(tag)
    g = golden(f"encoder_{tag}")
    xn = torch.from_numpy(g["x_norm"])
    tr = {}
    with torch.no_grad():
        oracle_R.encoder_forward(sd, xn, trace=tr)
    total_diff = 0
    for i in range(7):
        sf, df = tr["src_f"][i], tr["dst_f"][i]
        B, C, _, Ns = sf.shape
        idx, d2 = knn_graph_cm(df.reshape(B, C * 3, -1).to(dev), sf.reshape(B, C * 3, Ns).to(dev))
        torch.cuda.synchronize()
        ref = tr["knn_idx"][i]
        assert np.array_equal(ref.numpy(), g[f"knn_idx_{i}"].astype(np.int64))  # oracle == reference run
        bad, ndiff, nrows = _knn_rows_equivalent(idx, ref, df, sf)
        total_diff += ndiff
        assert bad == 0, f"layer {i}: {bad} rows differ beyond fp32 near-ties ({ndiff}/{nrows} rows differ)"
        assert ndiff <= max(2, nrows // 500), f"layer {i}: too many near-tie rows ({ndiff}/{nrows})"
        # returned squared distances are ascending and match fp64 to fp32 accuracy
        assert bool((d2[..., 1:] >= d2[..., :-1]).all())
    print(f"[{tag}] kNN rows differing by near-ties over all layers: {total_diff}")


def test_knn_edge_cases(dev):
    from livingscenes_b200.ops import knn_points

    # exact ties from duplicated points: lower index first; self first; Ns not a multiple of the tile
    g = torch.Generator().manual_seed(3)
    p = torch.randn(2, 300, 3, generator=g)
    p[:, 17] = p[:, 5]
    p[:, 250] = p[:, 5]
    d2, idx, _ = knn_points(p.to(dev), p.to(dev), K=16)
    from oracle.p3d_shim import knn_points as ref_knn

    _, ridx, _ = ref_knn(p, p, K=16)
    assert torch.equal(idx.cpu(), ridx)
    assert idx[0, 5, :3].tolist() == [5, 17, 250] and idx[0, 250, :3].tolist() == [5, 17, 250]
    # minimal source set (Ns == K) and high-dimensional features
    q = torch.randn(1, 40, 768, generator=g)
    s = torch.randn(1, 16, 768, generator=g)
    _, idx, _ = knn_points(q.to(dev), s.to(dev), K=16)
    _, ridx, _ = ref_knn(q, s, K=16)
    assert torch.equal(idx.cpu(), ridx)


# ------------------------------------------------------------------------------------------ tensor-core kNN
@pytest.mark.parametrize("tag", ["random", "shipped"])
def test_knn_tc_teacher_forced_equals_exact(tag, dev, oracle_R):
    """ls_knn_tc (tcgen05 candidate filter + exact re-rank) fed the oracle's layer inputs returns exactly what
    the brute-force ls_knn returns -- indices AND squared distances bit for bit -- and the oracle's graph."""
    from livingscenes_b200.ops import knn_graph_cm, knn_graph_cm_tc

    sd = state_dict_for(tag)
    g = golden(f"encoder_{tag}")
    xn = torch.from_numpy(g["x_norm"])
    tr = {}
    with torch.no_grad():
        oracle_R.encoder_forward(sd, xn, trace=tr)
    for i in range(7):
        sf, df = tr["src_f"][i], tr["dst_f"][i]
        B, C, _, Ns = sf.shape
        q, s = df.reshape(B, C * 3, -1).to(dev), sf.reshape(B, C * 3, Ns).to(dev)
        idx_e, d_e = knn_graph_cm(q, s)
        idx_t, d_t, nc = knn_graph_cm_tc(q, s)
        torch.cuda.synchronize()
        assert torch.equal(idx_e, idx_t), f"layer {i}: tensor-core graph differs from the brute-force graph"
        assert torch.equal(d_e, d_t), f"layer {i}: re-ranked distances differ bitwise"
        assert int(nc.min()) >= 16 and int(nc.max()) <= 64, f"layer {i}: candidate counts {int(nc.min())}..{int(nc.max())}"
        bad, ndiff, nrows = _knn_rows_equivalent(idx_t, tr["knn_idx"][i], df, sf)
        assert bad == 0
        print(f"[{tag}] layer {i}: candidates per query mean {float(nc.float().mean()):.1f} max {int(nc.max())}")


@pytest.mark.parametrize("B,D,Nq,Ns", [(2, 3, 300, 300), (3, 96, 1000, 1000), (1, 96, 512, 1024), (2, 192, 128, 512),
                                       (1, 8, 130, 2500), (2, 384, 40, 200), (1, 5, 16, 16)])
def test_knn_tc_shapes_and_ties(B, D, Nq, Ns, dev):
    """Ragged tiles (Nq, Ns not multiples of 128, D not a multiple of 8), exact ties from duplicated sources
    (lower index first), and a badly conditioned cloud (far from the origin: the filter's error budget exceeds
    the neighbour distances, lists overflow, the exact brute-force fallback takes over)."""
    from livingscenes_b200.ops import knn_graph_cm, knn_graph_cm_tc

    g = torch.Generator().manual_seed(B * 1000 + Nq)
    s = torch.randn(B, D, Ns, generator=g)
    if Ns > 40:
        s[:, :, 17] = s[:, :, 5]
        s[:, :, Ns - 3] = s[:, :, 5]
    q = s[:, :, :Nq].clone() if Nq <= Ns else torch.randn(B, D, Nq, generator=g)
    for shift in (0.0, 300.0):
        qq, ss = (q + shift).to(dev), (s + shift).to(dev)
        idx_e, d_e = knn_graph_cm(qq, ss)
        idx_t, d_t, nc = knn_graph_cm_tc(qq, ss)
        torch.cuda.synchronize()
        assert torch.equal(idx_e, idx_t), f"shift {shift}: graphs differ"
        assert torch.equal(d_e, d_t)
        if shift == 0.0 and Ns > 40 and Nq > 5:
            assert idx_t[0, 5, :3].tolist() == [5, 17, Ns - 3]
        print(f"shift {shift}: overflowed queries {int((nc < 0).sum())}/{nc.numel()}, max candidates {int(nc.max())}")


def test_knn_tc_overflow_fallback_is_exact(dev):
    """kappa_scale = 1e6 makes every candidate list overflow: the brute-force fallback must reproduce ls_knn."""
    from livingscenes_b200 import _lib
    from livingscenes_b200.ops import knn_graph_cm, knn_graph_cm_tc

    g = torch.Generator().manual_seed(11)
    s = torch.randn(2, 96, 700, generator=g).to(dev)
    q = torch.randn(2, 96, 260, generator=g).to(dev)
    idx_e, d_e = knn_graph_cm(q, s)
    try:
        _lib.set_knn_tensor_cores(True, 1e6)
        idx_t, d_t, nc = knn_graph_cm_tc(q, s)
        torch.cuda.synchronize()
    finally:
        _lib.set_knn_tensor_cores(True, 1.0)
    assert int((nc >= 0).sum()) == 0
    assert torch.equal(idx_e, idx_t) and torch.equal(d_e, d_t)


@pytest.mark.parametrize("tag", ["random", "shipped"])
@pytest.mark.parametrize("B,N", [(4, 1024), (2, 2048), (3, 1000)])
def test_encoder_knn_tc_equals_brute_force(tag, B, N, dev, oracle_R):
    """Whole encoder with the tensor-core graph vs the brute-force graph: identical indices at all 7 layers and
    bit-identical embeddings (the two paths feed the same arithmetic downstream)."""
    from livingscenes_b200 import _lib

    enc = _model(tag, dev).encoder
    x = oracle_R.synth_instances(B, N, 555 + N).to(dev)
    try:
        _lib.set_knn_tensor_cores(False, 1.0)
        a = enc.run(x, normalize=True, taps=True)
        _lib.set_knn_tensor_cores(True, 1.0)
        b = enc.run(x, normalize=True, taps=True)
        torch.cuda.synchronize()
    finally:
        _lib.set_knn_tensor_cores(True, 1.0)
    for i, (ia, ib) in enumerate(zip(a["knn_idx"], b["knn_idx"])):
        assert torch.equal(ia, ib), f"layer {i}"
    for k in ("z_so3", "z_inv", "scale", "center"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("kappa_scale", [0.25, 8.0, 64.0])
def test_knn_tc_error_budget_does_not_change_the_graph(kappa_scale, dev, oracle_R):
    """The filter's error budget only decides how many candidates are re-ranked and how many of them need exact
    distances (wider budget -> longer ambiguous runs): the graph must not depend on it.  0.25 is still ~10x the
    measured tensor-core error; 64 makes most neighbouring candidates 'ambiguous'."""
    from livingscenes_b200 import _lib
    from livingscenes_b200.ops import knn_graph_cm, knn_graph_cm_tc

    sd = state_dict_for("random")
    tr = {}
    with torch.no_grad():
        oracle_R.encode(sd, oracle_R.synth_instances(2, 1024, 77), trace=tr)
    for i in (0, 1, 3):
        sf, df = tr["src_f"][i], tr["dst_f"][i]
        B, C, _, Ns = sf.shape
        q, s = df.reshape(B, C * 3, -1).to(dev), sf.reshape(B, C * 3, Ns).to(dev)
        idx_e, d_e = knn_graph_cm(q, s)
        try:
            _lib.set_knn_tensor_cores(True, kappa_scale)
            idx_t, d_t, nc = knn_graph_cm_tc(q, s)
            torch.cuda.synchronize()
        finally:
            _lib.set_knn_tensor_cores(True, 1.0)
        assert torch.equal(idx_e, idx_t) and torch.equal(d_e, d_t), f"layer {i}, kappa_scale {kappa_scale}"


def test_encoder_overlap_and_budget_invariance(dev, oracle_R):
    """Side-stream overlap on/off and a wide error budget (hybrid re-rank computes exact distances for most
    candidates) give bit-identical graphs and embeddings; also under CUDA-graph capture + replay."""
    from livingscenes_b200 import _lib

    enc = _model("random", dev).encoder
    x = oracle_R.synth_instances(6, 1024, 4242).to(dev)
    ref = enc.run(x, normalize=True, taps=True)
    torch.cuda.synchronize()
    try:
        for overlap, ks in ((False, 1.0), (True, 32.0), (False, 32.0)):
            _lib.set_overlap(overlap)
            _lib.set_knn_tensor_cores(True, ks)
            r = enc.run(x, normalize=True, taps=True)
            torch.cuda.synchronize()
            for i, (ia, ib) in enumerate(zip(ref["knn_idx"], r["knn_idx"])):
                assert torch.equal(ia, ib), f"layer {i} overlap={overlap} kappa_scale={ks}"
            for k in ("z_so3", "z_inv", "scale", "center"):
                assert torch.equal(ref[k], r[k]), k
    finally:
        _lib.set_overlap(True)
        _lib.set_knn_tensor_cores(True, 1.0)
    # capture + replay with the side stream forked inside the capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        enc.run(x, normalize=True)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = enc.run(x, normalize=True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for k in ("z_so3", "z_inv", "scale", "center"):
        assert torch.equal(ref[k], out[k]), f"graph replay: {k}"


@pytest.mark.parametrize("N,n_out", [(20000, 1024), (9000, 300), (4000, 1024)])
def test_fps_start_index_and_large_clouds(N, n_out, dev):
    """ls_fps_ex: caller-chosen first index (random_start_point) and the large-cloud kernel (N > 8192) select
    exactly the oracle's points."""
    from livingscenes_b200.ops import farthest_point_sample
    from oracle.p3d_shim import sample_farthest_points as ref_fps

    g = torch.Generator().manual_seed(N)
    pts = torch.randn(2, N, 3, generator=g) * torch.tensor([1.0, 0.5, 2.0])
    start = torch.tensor([5, N - 3])
    _, ridx = ref_fps(pts, K=n_out, start_idx=start)
    idx, out = farthest_point_sample(pts.transpose(1, 2).to(dev), n_out, start.to(dev))
    torch.cuda.synchronize()
    assert torch.equal(idx.cpu(), ridx)
    assert torch.equal(out.cpu(), torch.gather(pts, 1, ridx[..., None].expand(-1, -1, 3)).transpose(1, 2))


def test_encode_fps_random_restarts(dev, oracle_R):
    """Shape_Prior.encode_fps(n_fps=3) (model_utils.py:199-215): ragged masked instances, three FPS restarts each
    with random first indices, codes averaged -- against the oracle fed the same start indices."""
    from oracle.p3d_shim import sample_farthest_points as ref_fps

    sd = state_dict_for("random")
    model = _model("random", dev)
    g = torch.Generator().manual_seed(9)
    B, Nmax, n_fps = 2, 3000, 3
    pc = torch.randn(B, 3, Nmax, generator=g) * 0.3
    mask = torch.zeros(B, 1, Nmax, dtype=torch.bool)
    mask[0, 0, :2500] = True
    mask[1, 0, 200:1700] = True
    torch.manual_seed(123)
    out = model.encode_fps(pc.to(dev), mask.to(dev), n_fps=n_fps)
    torch.cuda.synchronize()
    torch.manual_seed(123)
    ref = {k: [] for k in ("z_so3", "z_inv", "s", "t")}
    nv = mask.reshape(B, -1).sum(-1).repeat_interleave(n_fps).double()
    starts = torch.minimum((torch.rand(B * n_fps, dtype=torch.float64) * nv).long(), nv.long() - 1).reshape(B, n_fps)
    with torch.no_grad():
        for b in range(B):
            valid = pc[b][:, mask[b, 0]]
            sub, _ = ref_fps(valid.T[None].expand(n_fps, -1, -1).contiguous(), K=model.field_input_n, start_idx=starts[b])
            code = oracle_R.encode(sd, sub.transpose(1, 2).contiguous())
            for k in ref:
                ref[k].append(code[k].mean(0, keepdim=True))
    for k in ref:
        r = torch.cat(ref[k], 0)
        assert out[k].shape == r.shape, k
        assert relerr(out[k].cpu(), r) < TOL, k
    # deterministic path of the evals (n_fps = 1, start index 0)
    out1 = model.encode_fps(pc.to(dev), mask.to(dev))
    torch.cuda.synchronize()
    with torch.no_grad():
        subs = [ref_fps(pc[b][:, mask[b, 0]].T[None].contiguous(), K=model.field_input_n)[0] for b in range(B)]
        code1 = oracle_R.encode(sd, torch.cat(subs, 0).transpose(1, 2).contiguous())
    for k in ("z_so3", "z_inv", "s", "t"):
        assert relerr(out1[k].cpu(), code1[k]) < TOL, k
    # fewer valid points than the encoder samples: accepted like the reference -- pytorch3d's FPS selects every
    # point (FPS order) and pads the sample with zeros, which are encoded as points (model_utils.py:203-207)
    small = model.encode_fps(pc[:, :, :900].to(dev), torch.ones(B, 1, 900, dtype=torch.bool, device=dev))
    torch.cuda.synchronize()
    with torch.no_grad():
        subs = []
        for b in range(B):
            s900, _ = ref_fps(pc[b, :, :900].T[None].contiguous(), K=900)
            subs.append(torch.cat([s900, torch.zeros(1, model.field_input_n - 900, 3)], 1))
        code_s = oracle_R.encode(sd, torch.cat(subs, 0).transpose(1, 2).contiguous())
    for k in ("z_so3", "z_inv", "s", "t"):
        assert relerr(small[k].cpu(), code_s[k]) < TOL, k


@pytest.mark.parametrize("N,n_out", [(1024, 512), (2048, 1024), (1000, 77), (5000, 1024), (16, 16)])
def test_fps_bit_exact(N, n_out, dev, oracle_R):
    from livingscenes_b200.ops import farthest_point_sample
    from oracle.p3d_shim import sample_farthest_points as ref_fps

    x = oracle_R.synth_instances(3, N, 11 + N)
    idx, sub = farthest_point_sample(x.to(dev), n_out)
    rp, ridx = ref_fps(x.transpose(1, 2), K=n_out)
    assert torch.equal(idx.cpu(), ridx)
    assert torch.equal(sub.cpu(), rp.transpose(1, 2))


@pytest.mark.parametrize("N,n_out", [(1024, 512), (5000, 1024), (20000, 300)])
def test_fps_fma_switch_matches_the_contracted_oracle(N, n_out, dev, oracle_R):
    """ls_set_fps_fma(1): the FMA-contracted distance (what nvcc makes of pytorch3d's CUDA kernel) against the shim's
    emulation of it; both forms are bit-exact against their own oracle, and they do differ from each other somewhere."""
    from livingscenes_b200 import _lib
    from livingscenes_b200.ops import farthest_point_sample
    from oracle import p3d_shim

    x = oracle_R.synth_instances(2, N, 5 + N) * 37.3  # scaled: more mantissa bits in play
    try:
        _lib.set_fps_fma(True)
        p3d_shim.FPS_FMA = True
        idx, _ = farthest_point_sample(x.to(dev), n_out)
        _, ridx = p3d_shim.sample_farthest_points(x.transpose(1, 2), K=n_out)
        assert torch.equal(idx.cpu(), ridx)
    finally:
        _lib.set_fps_fma(False)
        p3d_shim.FPS_FMA = False
    idx0, _ = farthest_point_sample(x.to(dev), n_out)
    _, ridx0 = p3d_shim.sample_farthest_points(x.transpose(1, 2), K=n_out)
    assert torch.equal(idx0.cpu(), ridx0)


def test_fps_ties_lowest_index(dev):
    from livingscenes_b200.ops import farthest_point_sample

    x = torch.tensor([[[0.0, 1, -1, 0], [0, 0, 0, 1], [0, 0, 0, 0]]])  # points (0,0,0),(1,0,0),(-1,0,0),(0,1,0)
    idx, _ = farthest_point_sample(x.to(dev), 2)
    assert idx[0].tolist() == [0, 1]


# ------------------------------------------------------------------------------------------ VN-Linear GEMM
@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("B,Ci,Co,N", [(3, 32, 64, 1024), (5, 64, 256, 512), (7, 256, 2048, 32), (2, 512, 257, 32),
                                       (1, 257, 768, 1000), (4, 96, 40, 36)])
def test_vn_linear_fp32_accurate(tensor_cores, B, Ci, Co, N, dev):
    """VecLinear.forward through the C ABI: the FP32 SIMT GEMM and the tcgen05 3xTF32 GEMM both have to be
    fp32-accurate against a float64 reference -- single-pass TF32 would be ~1e-3."""
    import livingscenes_b200 as ls

    g = torch.Generator().manual_seed(B * 1000 + Ci)
    W = (torch.rand(Co, Ci, generator=g) * 2 - 1) / Ci ** 0.5
    v = torch.randn(B, Ci, 3, N, generator=g)
    out = ls.vn_linear(W.to(dev), v.to(dev), tensor_cores=tensor_cores)
    torch.cuda.synchronize()
    ref = torch.einsum("oc,bcan->boan", W.double(), v.double())
    assert out.shape == (B, Co, 3, N)
    err = float((out.cpu().double() - ref).abs().max() / ref.abs().max())
    # 3xTF32 on tcgen05: ~1e-6 at K <= 256, 4e-6 measured at K = 512 (the tensor core accumulates fp32 with
    # truncation, which adds a small K-proportional bias); single-pass TF32 would sit at ~5e-4.
    assert err < (1e-5 if tensor_cores else 2e-6), f"max-rel error {err:.2e}"


@pytest.mark.parametrize("B,Ci,Co,N", [(3, 32, 64, 1024), (37, 64, 384, 512), (300, 128, 768, 128), (7, 256, 2048, 32),
                                       (2, 512, 257, 32), (1, 257, 768, 1000), (4, 96, 40, 36), (256, 32, 128, 1024),
                                       (300, 256, 512, 32), (260, 512, 1100, 32)])
def test_gemm_persistent_equals_per_tile_kernel(B, Ci, Co, N, dev):
    """The persistent warp-specialised tcgen05 GEMM (bulk-TMA fed, double-buffered TMEM; many tiles per CTA at the
    larger sizes) issues the same MMAs in the same order as the round-1 one-tile-per-CTA kernel: bit-identical."""
    import livingscenes_b200 as ls
    from livingscenes_b200 import _lib

    g = torch.Generator().manual_seed(B * 1000 + Ci)
    W = ((torch.rand(Co, Ci, generator=g) * 2 - 1) / Ci ** 0.5).to(dev)
    v = torch.randn(B, Ci, 3, N, generator=g).to(dev)
    try:
        _lib.set_gemm_variant(1)
        a = ls.vn_linear(W, v, tensor_cores=True)
        _lib.set_gemm_variant(2)
        b = ls.vn_linear(W, v, tensor_cores=True)
        _lib.set_gemm_variant(3)
        c = ls.vn_linear(W, v, tensor_cores=True)
        torch.cuda.synchronize()
    finally:
        _lib.set_gemm_variant(3)
    assert torch.equal(a, b)
    ref = torch.einsum("oc,bcan->boan", W.double().cpu(), v.double().cpu())
    assert float((b.cpu().double() - ref).abs().max() / ref.abs().max()) < 1e-5
    # variant 3 (activations in tensor memory, operand roles swapped, 256-row weight tiles for K >= 128): the same three
    # TF32 products per k-step, fp32-accurate like the others
    assert c.shape == b.shape
    assert float((c.cpu().double() - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float((c - b).abs().max() / b.abs().max()) < 4e-6


def test_encoder_wave_schedule_invariance(dev, oracle_R):
    """The {table GEMM -> EdgeConv} wave schedule (L2-resident gather tables, two alternating slots, side stream)
    does not change a single bit: whole-batch launches, tiny waves, GEMM variant 1, with and without overlap."""
    from livingscenes_b200 import _lib

    enc = _model("random", dev).encoder
    x = oracle_R.synth_instances(23, 1024, 777).to(dev)
    try:
        # variants 1 and 2 issue identical MMAs (bit-identical to each other); variant 3 swaps the operand roles and is
        # compared with itself
        for ref_variant, cases in ((2, ((28, True, 2), (6, True, 2), (6, False, 2), (12, True, 1), (1, True, 2))),
                                   (3, ((28, True, 3), (6, False, 3), (1, True, 3)))):
            _lib.set_wave_bytes(0)
            _lib.set_overlap(True)
            _lib.set_gemm_variant(ref_variant)
            ref = enc.run(x, normalize=True, taps=True)
            torch.cuda.synchronize()
            for wave_mb, overlap, variant in cases:
                _lib.set_wave_bytes(wave_mb << 20)
                _lib.set_overlap(overlap)
                _lib.set_gemm_variant(variant)
                for _ in range(2):
                    r = enc.run(x, normalize=True, taps=True)
                torch.cuda.synchronize()
                for i, (ia, ib) in enumerate(zip(ref["knn_idx"], r["knn_idx"])):
                    assert torch.equal(ia, ib), f"layer {i} wave_mb={wave_mb}"
                for i, (fa, fb) in enumerate(zip(ref["feat"], r["feat"])):
                    assert torch.equal(fa, fb), f"features of layer {i} wave_mb={wave_mb} overlap={overlap} variant={variant}"
                for k in ("z_so3", "z_inv", "scale", "center"):
                    assert torch.equal(ref[k], r[k]), k
    finally:
        _lib.set_wave_bytes(0)
        _lib.set_overlap(True)
        _lib.set_gemm_variant(3)


# ------------------------------------------------------------------------------------------ encoder
@pytest.mark.parametrize("tag", ["random", "shipped"])
def test_encoder_teacher_forced_matches_golden(tag, dev):
    """VecDGCNN_att.forward with the reference's graph forced: every layer's features and the head."""
    g = golden(f"encoder_{tag}")
    m = _model(tag, dev)
    xn = torch.from_numpy(g["x_norm"]).to(dev)
    knn = [torch.from_numpy(g[f"knn_idx_{i}"].astype(np.int64)) for i in range(7)]
    fps = [torch.from_numpy(g[f"fps_idx_{i}"].astype(np.int64)) for i in range(3)]
    r = m.encoder.run(xn, taps=True, force_knn_idx=knn, force_fps_idx=fps)
    torch.cuda.synchronize()
    for i in range(7):
        assert torch.equal(r["knn_idx"][i].cpu(), knn[i])
        e = relerr(r["feat"][i][..., ::16], g[f"feat_{i}"])
        assert e < TOL, f"layer {i} features: {e:.2e}"
    for k in ("center", "scale", "z_so3", "z_inv"):
        e = relerr(r[k].reshape(g[k].shape), g[k])
        assert e < TOL, f"{k}: {e:.2e}"


@pytest.mark.parametrize("tag,name", [("random", "encoder_random"), ("shipped", "encoder_shipped"),
                                      ("random", "encoder_random_n2048"), ("shipped", "encoder_shipped_n2048")])
def test_encoder_free_running_matches_golden(tag, name, dev):
    """No forcing: the CUDA path builds its own graph (FPS + kNN) and must land on the reference's
    embedding; indices are compared exactly and the differing rows are reported."""
    g = golden(name)
    m = _model(tag, dev)
    xn = torch.from_numpy(g["x_norm"]).to(dev)
    r = m.encoder.run(xn, taps=True)
    out = m.encoder(xn)
    torch.cuda.synchronize()
    assert len(out) == 4 and out[0].shape == (xn.shape[0], 1, 3)
    for j in range(3):
        assert torch.equal(r["fps_idx"][j].cpu(), torch.from_numpy(g[f"fps_idx_{j}"].astype(np.int64))), f"FPS {j}"
    n_diff = 0
    for i in range(7):
        ref = torch.from_numpy(g[f"knn_idx_{i}"].astype(np.int64))
        rows = (r["knn_idx"][i].cpu() != ref).any(-1)
        sets = (r["knn_idx"][i].cpu().sort(-1)[0] != ref.sort(-1)[0]).any(-1)
        n_diff += int(rows.sum())
        assert int(sets.sum()) <= max(1, ref.shape[0] * ref.shape[1] // 1000), \
            f"layer {i}: {int(sets.sum())} neighbour SETS differ"
    print(f"[{name}] kNN rows with any ordered difference: {n_diff}")
    for k in ("center", "scale", "z_so3", "z_inv"):
        e = relerr(r[k].reshape(g[k].shape), g[k])
        assert e < TOL, f"{k}: {e:.2e}"
    assert torch.equal(out[2], r["z_so3"])  # forward() is deterministic and equals run()


@pytest.mark.parametrize("tag,name", [("random", "encoder_random"), ("shipped", "encoder_shipped"),
                                      ("shipped", "encoder_shipped_n2048")])
def test_shape_prior_encode_matches_golden(tag, name, dev):
    g = golden(name)
    m = _model(tag, dev)
    x = torch.from_numpy(g["x"]).to(dev)
    r = m.encoder.run(x, normalize=True, taps=True)
    code = m.encode(x)
    torch.cuda.synchronize()
    assert relerr(r["scale0"], g["scale0"]) < 1e-5
    assert relerr(r["x_norm"], g["x_norm"]) < 1e-5
    assert code["t"].shape == (x.shape[0], 1, 3) and code["s"].shape == (x.shape[0],)
    for k, gk in (("z_so3", "enc_z_so3"), ("z_inv", "enc_z_inv"), ("s", "enc_s"), ("t", "enc_t")):
        e = relerr(code[k], g[gk])
        assert e < TOL, f"{k}: {e:.2e}"


def test_encoder_against_oracle_fresh_inputs(dev, oracle_R):
    """Same seeded inputs through the oracle restatement (CPU) and the CUDA path; B=2, N=512."""
    sd = state_dict_for("random")
    m = _model("random", dev)
    x = oracle_R.synth_instances(2, 512, 4321)
    with torch.no_grad():
        ref = oracle_R.encode(sd, x)
    code = m.encode(x.to(dev))
    for k in ref:
        e = relerr(code[k], ref[k])
        assert e < TOL, f"{k}: {e:.2e}"


@pytest.mark.parametrize("B,N", [(1, 512), (2, 1000), (5, 640), (1, 4096)])
def test_encoder_other_sizes_against_oracle(B, N, dev, oracle_R):
    """Sizes off the beaten path: N = 1000 / 640 give point counts that are not multiples of the tile sizes
    (and 3*N/32 columns that force the FP32 SIMT GEMM fallback for the deep layers), B = 1 is the
    3RScan per-instance call pattern, N = 4096 exercises 16 source tiles per query."""
    sd = state_dict_for("random")
    m = _model("random", dev)
    x = oracle_R.synth_instances(B, N, 900 + N)
    tr = {}
    with torch.no_grad():
        ref = oracle_R.encode(sd, x, trace=tr)
    r = m.encoder.run(x.to(dev), normalize=True, taps=True)
    torch.cuda.synchronize()
    for j in range(3):
        assert torch.equal(r["fps_idx"][j].cpu(), tr["fps_idx"][j]), f"FPS {j}"
    # free-running graphs may differ at fp32 near-ties; the embedding must agree with the oracle driven by the
    # CUDA graph, and (loosely) with the oracle's own run
    with torch.no_grad():
        c, s_, zs, zi = oracle_R.encoder_forward(sd, r["x_norm"].cpu(), force={"knn_idx": [t.cpu() for t in r["knn_idx"]],
                                                                              "fps_idx": [t.cpu() for t in r["fps_idx"]]})
    assert relerr(r["z_so3"], zs) < TOL and relerr(r["z_inv"], zi) < TOL
    n_diff = sum(int((a.cpu() != b_).any(-1).sum()) for a, b_ in zip(r["knn_idx"], tr["knn_idx"]))
    n_rows = sum(a.shape[0] * a.shape[1] for a in tr["knn_idx"])
    assert n_diff <= max(2, n_rows // 2000), f"{n_diff}/{n_rows} kNN rows differ from the oracle"
    assert relerr(r["z_inv"], ref["z_inv"]) < (TOL if n_diff == 0 else 5e-2)


def test_batch_consistency_and_equivariance_full_size(dev, oracle_R):
    """BASELINE config 2 size (B=256, N=1024): (a) an instance encodes identically alone and inside the
    batch (bit-exact: no cross-instance leakage through the flattened GEMM columns); (b) the
    equivariance property the reference's __main__ prints (vec_dgcnn_atten.py:279-319): rotating and
    scaling the input gives z_so3 R^T, scale*s, unchanged z_inv."""
    m = _model("random", dev)
    x = oracle_R.synth_instances(256, 1024, 1235).to(dev)
    xc = x - x.mean(-1, keepdim=True)
    full = m.encoder.run(xc)
    for b in (0, 100, 255):
        one = m.encoder.run(xc[b:b + 1].contiguous())
        for k in ("center", "scale", "z_so3", "z_inv"):
            assert torch.equal(one[k][0], full[k][b]), f"instance {b} {k} differs inside the batch"
    Rm = oracle_R.random_rotations(256, 5).to(dev)
    s = (0.5 + torch.rand(256, generator=torch.Generator().manual_seed(1))).to(dev)
    xr = torch.einsum("bij,bjn->bin", Rm, xc) * s[:, None, None]
    rot = m.encoder.run(xr.contiguous())
    torch.cuda.synchronize()
    z_rot = torch.einsum("bcj,bij->bci", full["z_so3"], Rm)
    # The rotated cloud rounds differently in fp32, so a few instances take a different (near-tie)
    # FPS / kNN decision and drift more; judge the distribution over the 256 instances.
    def per_instance(a, b):
        a, b = a.reshape(256, -1).double(), b.reshape(256, -1).double()
        return ((a - b).abs().amax(1) / b.abs().amax(1)).cpu()
    for name, e in (("z_so3", per_instance(rot["z_so3"], z_rot)), ("z_inv", per_instance(rot["z_inv"], full["z_inv"])),
                    ("scale", per_instance(rot["scale"], full["scale"] * s))):
        print(f"equivariance {name}: median {float(e.median()):.2e}  p95 {float(e.quantile(0.95)):.2e}  max {float(e.max()):.2e}")
        assert float(e.median()) < 1e-3, name
        assert float(e.quantile(0.95)) < 2e-2, name
        assert float(e.max()) < 0.5, name


# ------------------------------------------------------------------------------------------ solvers
def test_matchers_match_golden(dev):
    import livingscenes_b200 as ls

    g = golden("solver_cases")
    for ci in range(int(g["n_match_cases"])):
        z0, z1 = torch.from_numpy(g[f"m{ci}_z0"]).to(dev), torch.from_numpy(g[f"m{ci}_z1"]).to(dev)
        r = ls.sequential_matcher(z0, z1)
        assert r["matches0"].dtype == torch.int64
        assert np.array_equal(r["matches0"].cpu().numpy(), g[f"m{ci}_seq0"]), f"case {ci} matches0"
        assert np.array_equal(r["matches1"].cpu().numpy(), g[f"m{ci}_seq1"]), f"case {ci} matches1"
        rn = ls.nn_matcher(z0.T[None], z1.T[None])
        assert np.array_equal(rn["matches0"].reshape(-1).cpu().numpy(), g[f"m{ci}_nn0"]), f"case {ci} nn0"
        assert np.array_equal(rn["matches1"].reshape(-1).cpu().numpy(), g[f"m{ci}_nn1"]), f"case {ci} nn1"


def test_matcher_batched_equals_single(dev):
    import livingscenes_b200 as ls

    g = torch.Generator().manual_seed(8)
    sizes0, sizes1 = [5, 32, 1, 17] * 15, [7, 32, 4, 9] * 15  # 60 pairs: more than one launch
    z0 = torch.randn(sum(sizes0), 256, generator=g).to(dev)
    z1 = torch.randn(sum(sizes1), 256, generator=g).to(dev)
    rb = ls.sequential_matcher_batched(z0, z1, sizes0, sizes1)
    o0 = o1 = 0
    for n, k in zip(sizes0, sizes1):
        r = ls.sequential_matcher(z0[o0:o0 + n], z1[o1:o1 + k])
        assert torch.equal(rb["matches0"][o0:o0 + n], r["matches0"])
        assert torch.equal(rb["matches1"][o1:o1 + k], r["matches1"])
        o0, o1 = o0 + n, o1 + k


def test_matcher_against_oracle_random(dev, oracle_R):
    import livingscenes_b200 as ls

    g = torch.Generator().manual_seed(21)
    for n, k in [(32, 32), (3, 40), (64, 50), (128, 128)]:
        z0 = torch.randn(n, 256, generator=g)
        z1 = torch.cat([z0[torch.randperm(n, generator=g)][:min(n, k)] + 0.3 * torch.randn(min(n, k), 256, generator=g),
                        torch.randn(max(0, k - n), 256, generator=g)])
        ref = oracle_R.sequential_match(z0, z1)
        r = ls.sequential_matcher(z0.to(dev), z1.to(dev))
        assert torch.equal(r["matches0"].cpu(), ref["matches0"]) and torch.equal(r["matches1"].cpu(), ref["matches1"])
        refn = oracle_R.mutual_nn_match(z0.T[None], z1.T[None])
        rn = ls.nn_matcher(z0.T[None].to(dev), z1.T[None].to(dev))
        assert torch.equal(rn["matches0"].cpu(), refn["matches0"]) and torch.equal(rn["matches1"].cpu(), refn["matches1"])


def test_kabsch_matches_golden(dev):
    import livingscenes_b200 as ls

    g = golden("solver_cases")
    x1, x2, w = (torch.from_numpy(g[k]).to(dev) for k in ("k_x1", "k_x2", "k_w"))
    R, t, res, flag = ls.kabsch_transformation_estimation(x1, x2)
    assert R.shape == (6, 3, 3) and t.shape == (6, 3, 1) and res.shape == (6, 256) and flag is False
    assert float((R.cpu() - torch.from_numpy(g["k_R"])).abs().max()) < TOL
    assert float((t.cpu() - torch.from_numpy(g["k_t"])).abs().max()) < TOL
    assert float((res.cpu() - torch.from_numpy(g["k_res"])).abs().max()) < TOL
    Rw, tw, resw, _ = ls.kabsch_transformation_estimation(x1, x2, weights=w)
    assert float((Rw.cpu() - torch.from_numpy(g["k_Rw"])).abs().max()) < TOL
    assert float((tw.cpu() - torch.from_numpy(g["k_tw"])).abs().max()) < TOL
    assert float((resw.cpu() - torch.from_numpy(g["k_resw"])).abs().max()) < TOL
    assert torch.allclose(torch.det(R.cpu()), torch.ones(6), atol=1e-5)


@pytest.mark.parametrize("tag", ["random", "shipped"])
def test_pair_pipeline_matches_golden(tag, dev):
    """BASELINE config 3 shape (N=2048 pairs): encode both sets, sequential match, Kabsch per pair."""
    import livingscenes_b200 as ls

    g = golden(f"pair_{tag}")
    m = _model(tag, dev)
    solver = ls.More_Solver(m)
    out = solver.solve_scene_pair(torch.from_numpy(g["xa"]).to(dev), torch.from_numpy(g["xb"]).to(dev))
    torch.cuda.synchronize()
    assert np.array_equal(out["matches"]["matches0"].cpu().numpy(), g["matches0"])
    assert np.array_equal(out["matches"]["matches1"].cpu().numpy(), g["matches1"])
    if tag == "shipped":
        assert relerr(out["ref_codes"]["z_inv"], g["za_inv"]) < TOL
        assert relerr(out["rescan_codes"]["z_so3"], g["zb_so3"]) < TOL
    else:
        # untrained weights amplify fp32 near-tie graph flips: require the embedding to equal the ORACLE's
        # when the oracle is driven with the graph the CUDA path actually built (and bound the raw drift)
        from oracle import restatement as R

        assert relerr(out["ref_codes"]["z_inv"], g["za_inv"]) < 2e-2
        xa = torch.from_numpy(g["xa"]).to(dev)
        r = m.encoder.run(xa, normalize=True, taps=True)
        with torch.no_grad():
            c, s_, zs, zi = R.encoder_forward(state_dict_for(tag), r["x_norm"].cpu(),
                                              force={"knn_idx": [t.cpu() for t in r["knn_idx"]],
                                                     "fps_idx": [t.cpu() for t in r["fps_idx"]]})
        assert relerr(r["z_so3"], zs) < TOL and relerr(r["z_inv"], zi) < TOL
    # kernel-level pose parity on the reference's own embeddings (well-posed for both weight sets)
    ca = {"z_so3": torch.from_numpy(g["za_so3"]).to(dev), "t": torch.from_numpy(g["ta"]).to(dev)}
    cb = {"z_so3": torch.from_numpy(g["zb_so3"]).to(dev), "t": torch.from_numpy(g["tb"]).to(dev)}
    R, t, res = ls.kabsch_from_codes(ca, cb, torch.from_numpy(g["matches0"]).to(dev))
    tol = TOL if tag == "shipped" else 5e-3  # random weights: ill-conditioned 3x3 problem (see make_golden.py)
    assert float((R.cpu() - torch.from_numpy(g["R"])).abs().max()) < tol
    assert float((t.cpu() - torch.from_numpy(g["t"])).abs().max()) < tol * max(1.0, float(np.abs(g["t"]).max()))
    if tag == "shipped":  # end to end
        assert float((out["R"].cpu() - torch.from_numpy(g["R"])).abs().max()) < 1e-3
        assert relerr(out["t"], g["t"]) < 1e-3
    nn = solver._solve_object_matching(out["ref_codes"], out["rescan_codes"], "nn")
    assert np.array_equal(nn["matches0"].reshape(-1).cpu().numpy(), g["nn_matches0"].reshape(-1))


# ------------------------------------------------------------------------------------------ SDF
@pytest.mark.parametrize("tag", ["random", "shipped"])
def test_sdf_matches_golden(tag, dev):
    g = golden(f"sdf_{tag}")
    m = _model(tag, dev)
    code = {k: torch.from_numpy(g[k]).to(dev) for k in ("z_so3", "z_inv", "s", "t")}
    q = torch.from_numpy(g["query"]).to(dev)
    sdf = m.decoder(q, None, code, return_sdf=True)
    occ = m.decoder(q, None, code)
    torch.cuda.synchronize()
    assert sdf.shape == q.shape[:2]
    assert float((sdf.cpu() - torch.from_numpy(g["sdf"])).abs().max()) < TOL
    assert torch.equal(occ.logits, -sdf)


def test_sdf_chunking_is_consistent(dev, oracle_R):
    """More columns than one pass holds (B*M > 131072): chunked result equals per-instance calls."""
    m = _model("random", dev)
    g = golden("sdf_random")
    code = {k: torch.from_numpy(g[k]).to(dev) for k in ("z_so3", "z_inv", "s", "t")}
    gen = torch.Generator().manual_seed(5)
    q = ((torch.rand(2, 70001, 3, generator=gen) - 0.5) * 1.1).to(dev) * code["s"][:, None, None] + code["t"]
    both = m.decoder(q, None, code, return_sdf=True)
    for b in range(2):
        one = m.decoder(q[b:b + 1].contiguous(), None, {k: v[b:b + 1] for k, v in code.items()}, return_sdf=True)
        assert float((one[0] - both[b]).abs().max()) < 1e-6
    with torch.no_grad():
        ref = oracle_R.sdf_decode(state_dict_for("random"), q[:, :4096].cpu(), {k: v.cpu() for k, v in code.items()})
    assert float((both[:, :4096].cpu() - ref).abs().max()) < TOL


# ------------------------------------------------------------------------------------------ ICP refinement
@pytest.mark.parametrize("N,M,with_init", [(1024, 1024, True), (700, 1500, False), (1024, 1024, False)])
def test_icp_matches_oracle(N, M, with_init, dev):
    """ls_icp against the restated pytorch3d iterative_closest_point (more_solver.py:182-187)."""
    import math

    from livingscenes_b200.ops import SimilarityTransform, iterative_closest_point
    from oracle import p3d_shim

    g = torch.Generator().manual_seed(N + M)
    B = 3
    Y = torch.randn(B, M, 3, generator=g) * torch.tensor([1.0, 0.6, 0.8])
    ang = torch.tensor([0.12, -0.2, 0.05])
    Rz = torch.stack([torch.tensor([[math.cos(a), -math.sin(a), 0.0], [math.sin(a), math.cos(a), 0.0], [0.0, 0.0, 1.0]])
                      for a in ang.tolist()])
    X = torch.bmm(Y[:, :N] if N <= M else Y.repeat(1, 2, 1)[:, :N], Rz) + torch.tensor([0.05, -0.03, 0.02])
    X = X + 0.002 * torch.randn(B, N, 3, generator=g)
    init = None
    if with_init:
        init = SimilarityTransform(Rz.transpose(1, 2) @ torch.eye(3), torch.tensor([[-0.04, 0.02, -0.01]]).repeat(B, 1),
                                   torch.ones(B))
    ref = p3d_shim.iterative_closest_point(X, Y, init_transform=init)
    dinit = None if init is None else SimilarityTransform(*(t.to(dev) for t in init))
    sol = iterative_closest_point(X.to(dev), Y.to(dev), init_transform=dinit)
    torch.cuda.synchronize()
    assert sol.converged == ref.converged
    assert float((sol.RTs.R.cpu() - ref.RTs.R).abs().max()) < TOL
    assert float((sol.RTs.T.cpu() - ref.RTs.T).abs().max()) < TOL
    assert float((sol.Xt.cpu() - ref.Xt).abs().max()) < 5 * TOL
    assert float((sol.rmse.cpu() - ref.rmse).abs().max()) < TOL
    assert float((torch.det(sol.RTs.R.cpu()) - 1).abs().max()) < 1e-5


def test_pairwise_registration_with_icp(dev, oracle_R):
    """More_Solver._solve_pairwise_registration (more_solver.py:95-116,182-189): FPS -> encode -> Kabsch on the
    equivariant codes -> ICP, against the oracle chain on the same clouds."""
    import livingscenes_b200 as ls
    from oracle import p3d_shim

    sd = state_dict_for("random")
    model = _model("random", dev)
    solver = ls.More_Solver(model)
    pc1 = oracle_R.synth_instances(1, 1500, 31).transpose(1, 2).contiguous()  # [1,N,3]
    Rg = oracle_R.random_rotations(1, 32)[0]
    pc2 = (pc1 @ Rg.T) + torch.tensor([0.2, -0.1, 0.3])
    R, t = solver._solve_pairwise_registration(pc1.to(dev), pc2.to(dev))
    torch.cuda.synchronize()
    with torch.no_grad():
        s1, _ = p3d_shim.sample_farthest_points(pc1, K=1024)
        s2, _ = p3d_shim.sample_farthest_points(pc2, K=1024)
        c1, c2 = oracle_R.encode(sd, s1.transpose(1, 2).contiguous()), oracle_R.encode(sd, s2.transpose(1, 2).contiguous())
        Rk, tk, _ = oracle_R.kabsch(c1["z_so3"] + c1["t"], c2["z_so3"] + c2["t"])
        ref = p3d_shim.iterative_closest_point(
            s1, s2, init_transform=p3d_shim.SimilarityTransform(Rk.transpose(-1, -2), tk.squeeze(2), torch.ones(1)))
    Rr, tr = ref.RTs.R.transpose(-1, -2), ref.RTs.T.unsqueeze(2)
    assert float((R.cpu() - Rr).abs().max()) < 5 * TOL and float((t.cpu() - tr).abs().max()) < 5 * TOL
    # the refined pose maps pc1 onto pc2
    err = ((pc1 @ R.cpu()[0].T + t.cpu()[0].T) - pc2).norm(dim=-1).max()
    assert float(err) < 1e-3


def test_fps_masked_batch_equals_per_instance(dev):
    """ls_fps_masked: one launch over a ragged masked batch == FPS of pc[:, mask] per instance (oracle), incl. a
    large instance and scattered masks."""
    from livingscenes_b200.ops import farthest_point_sample_masked
    from oracle.p3d_shim import sample_farthest_points as ref_fps

    g = torch.Generator().manual_seed(21)
    B, Nmax, n_out = 3, 12000, 1024
    pc = torch.randn(B, 3, Nmax, generator=g)
    mask = torch.zeros(B, Nmax, dtype=torch.bool)
    mask[0, :11000] = True
    mask[1] = torch.rand(Nmax, generator=g) < 0.3
    mask[2, 5000:6500] = True
    start = torch.tensor([0, 17, 1499])
    out, nv = farthest_point_sample_masked(pc.to(dev), mask.to(dev), n_out, start.to(dev))
    torch.cuda.synchronize()
    assert nv.cpu().tolist() == mask.sum(-1).tolist()
    for b in range(B):
        valid = pc[b][:, mask[b]].T[None].contiguous()
        pts, _ = ref_fps(valid, K=n_out, start_idx=start[b:b + 1])
        assert torch.equal(out[b].cpu(), pts[0].T), f"instance {b}"
