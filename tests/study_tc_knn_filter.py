"""CPU emulation of the tensor-core kNN prefilter (design study for k_knn_tc):
approximate ranking value  dt = |s|^2 - 2 q.s  with q.s from a 3xTF32 product, threshold from group minima,
candidate set {dt <= tau + 2E}, exact fp32 direct-form re-rank.  Reports candidate counts and error margins."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import restatement as R

def tf32_rn(x):
    u = x.view(np.uint32)
    h = ((u + 0x1000) & 0xffffe000).astype(np.uint32)
    return h.view(np.float32)
def tf32_trunc(x):
    return (x.view(np.uint32) & 0xffffe000).astype(np.uint32).view(np.float32)

def main():
    sd = torch.load(os.path.join(ROOT, "livingscenes_b200/_weights/shipped_fp32.pt"), map_location="cpu", weights_only=True)
    if len(sys.argv) > 1 and sys.argv[1] == "random":
        sd = R.random_state_dict(0)
    x = R.synth_instances(3, 1024, 4321)
    tr = {}
    with torch.no_grad():
        R.encode(sd, x, trace=tr)
    for i in range(5):
        src = tr["src_f"][i]; dst = tr["dst_f"][i]
        B, C, _, Ns = src.shape; Nd = dst.shape[-1]; D = 3 * C
        for center in (False, True):
            S = src.reshape(B, D, Ns).numpy().astype(np.float32)
            Q = dst.reshape(B, D, Nd).numpy().astype(np.float32)
            if center:
                mu = S.mean(-1, keepdims=True).astype(np.float32)
                S = (S - mu).astype(np.float32); Q = (Q - mu).astype(np.float32)
            rat, cands, errs, cands2 = [], [], [], []
            for b in range(B):
                s, q = S[b], Q[b]                       # [D][N]
                ns = np.zeros(Ns, np.float32)
                for d in range(D): ns = (s[d] * s[d] + ns).astype(np.float32)
                nq = (q.astype(np.float64) ** 2).sum(0)
                # exact-ish reference distance (float64 of the ORIGINAL fp32 inputs is translation invariant up to rounding)
                dex = ((q.T[:, None, :].astype(np.float64) - s.T[None, :, :].astype(np.float64)) ** 2).sum(-1)  # Nd x Ns
                sh, qh = tf32_rn(s), tf32_rn(q)
                sl, ql = tf32_trunc((s - sh).astype(np.float32)), tf32_trunc((q - qh).astype(np.float32))
                dot = (qh.T.astype(np.float64) @ sh.astype(np.float64) + qh.T.astype(np.float64) @ sl.astype(np.float64)
                       + ql.T.astype(np.float64) @ sh.astype(np.float64)).astype(np.float32)
                dt = (ns[None, :] - 2 * dot).astype(np.float32)
                err = np.abs(dt.astype(np.float64) + nq[:, None] - dex)
                nsmax = ns.max()
                scale = nq[:, None] + nsmax
                errs.append((err / scale).max())
                d16 = np.sort(dex, 1)[:, 15]
                rat.append(np.median(d16 / (nq + nsmax)))
                kappa = D * 2.0 ** -23 + 2.0 ** -16
                E = kappa * (nq + nsmax)
                # threshold: 16th smallest of 32 group minima over the first 512 sources (groups = s mod 32)
                sub = dt[:, :min(512, Ns)]
                gm = sub.reshape(Nd, -1, 32).min(1)
                tau = np.sort(gm, 1)[:, 15]
                cnt = (dt <= (tau + 2 * E)[:, None]).sum(1)
                cands.append(cnt)
                gm2 = dt.reshape(Nd, -1, 32).min(1)
                tau2 = np.sort(gm2, 1)[:, 15]
                cands2.append((dt <= (tau2 + 2 * E)[:, None]).sum(1))
                # inclusion check
                top = np.argsort(dex, 1, kind="stable")[:, :16]
                inc = np.take_along_axis(dt, top, 1) <= (tau + 2 * E)[:, None]
                assert inc.all(), "exact top-16 not inside the candidate set"
            c = np.concatenate(cands); c2 = np.concatenate(cands2)
            print(f"layer {i} D={D} Ns={Ns} Nd={Nd} center={center}: median d16/(nq+nsmax)={np.median(rat):.3e} "
                  f"max err/(nq+nsmax)={max(errs):.2e} (kappa {kappa:.2e}) cand(512-sample) mean {c.mean():.1f} p99 {np.percentile(c,99):.0f} max {c.max()} | cand(full) mean {c2.mean():.1f} max {c2.max()}")
main()
