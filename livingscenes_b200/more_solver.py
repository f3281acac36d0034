"""More_Solver -- the inference orchestration of lib_more/more_solver.py on the CUDA hot path.

  _solve_object_matching        more_solver.py:71-93   all five methods (nn, sinkhorn, sequential, sim3_seq, eq_seq)
  _solve_pairwise_registration  more_solver.py:95-189  FPS -> encode -> Kabsch on z_so3 + t -> [optim: SE(3) refinement
                                                       on the SDF + Sinkhorn loss, lib `optim.py`] -> ICP
  _optimize_code                more_solver.py:191-228 Adam on (z_inv, t, z_so3) against the SDF of the observed points
  _transform_latent             more_solver.py:230-244
  _solve_end2end                more_solver.py:246-290 batched: ONE ragged FPS launch per scan, ONE encode per scan, one
                                                       match, ONE Kabsch launch and ONE ICP launch for all matched
                                                       pairs (the reference re-runs FPS + a B=1 encode per matched pair;
                                                       both are deterministic, so the codes are the same)
  _mesh_from_latent / _mesh_from_pc  more_solver.py:37-69  through the GPU MISE extractor (mesh_extractor.py)
  solve_scene_pair              encode -> match -> pose for two pre-sampled instance sets (bench.py's step)
"""
from __future__ import annotations

import torch

from .matcher_new import eq_seq_matcher, nn_matcher, sequential_matcher, sim3_seq_matcher, sinkhorn_matcher
from .ops import SimilarityTransform, farthest_point_sample, farthest_point_sample_masked, iterative_closest_point
from .pose_estimation import kabsch_from_codes, kabsch_transformation_estimation

DEFAULT_CFG = {
    "shape_priors": {"n_input_point": 1024},
    "fps": {"n_init": 1, "random_start": False},
    # configs/more_3rscan.yaml
    "registration": {"optim": True, "step_size": {"so3": 0.05}, "n_steps": 400, "early_stop_threshold": 10},
    "mesh_extractor": {"threshold": 0.5, "resolution0": 32, "upsampling_steps": 2, "sample": False,
                       "simplify_nfaces": 5000, "points_batch_size": 10000, "refinement_step": 0},
}


def Rt_to_SE3(R, t):
    """lib_math/torch_se3.py: [b,3,3], [b,3,1] -> [b,3,4]."""
    return torch.cat([R, t], dim=-1)


def se3_inverse(g):
    R, t = g[..., :3], g[..., 3:]
    Rt = R.transpose(-1, -2)
    return torch.cat([Rt, -Rt @ t], dim=-1)


def se3_transform(g, pts):
    """g [b,3,4], pts [b,n,3] -> [b,n,3]."""
    return pts @ g[..., :3].transpose(-1, -2) + g[..., 3].unsqueeze(-2)


class More_Solver:
    def __init__(self, model, cfg=None) -> None:
        """``model`` is a ``livingscenes_b200.Shape_Prior``; ``cfg`` follows configs/more_3rscan.yaml (missing sections
        take that file's values)."""
        self.model = model
        self.cfg = {k: dict(v) for k, v in DEFAULT_CFG.items()}
        for k, v in (cfg or {}).items():
            if isinstance(v, dict) and k in self.cfg:
                self.cfg[k].update(v)
            else:
                self.cfg[k] = v
        if self.cfg["fps"].get("n_init", 1) != 1:
            raise NotImplementedError("fps.n_init > 1 in the solver is not built (Shape_Prior.encode_fps has n_fps)")
        self._mesh_extractor = None

    # ------------------------------------------------------------------ matching
    def _solve_object_matching(self, src_codes, tgt_codes, method):
        inv_src = src_codes["z_inv"].detach()
        inv_tgt = tgt_codes["z_inv"].detach()
        if method == "nn":
            return nn_matcher(inv_src.T[None], inv_tgt.T[None])
        if method == "sinkhorn":
            return sinkhorn_matcher(inv_src.T[None], inv_tgt.T[None])
        if method == "sequential":
            return sequential_matcher(inv_src, inv_tgt)
        if method == "sim3_seq":
            return sim3_seq_matcher(src_codes, tgt_codes)
        if method == "eq_seq":
            return eq_seq_matcher(src_codes, tgt_codes)
        raise ValueError(f"unknown matching method {method!r}")

    # ------------------------------------------------------------------ registration
    def _solve_pairwise_registration(self, pc1_full, pc2_full, optim=False, icp=True):
        """pc1 [1,N,3], pc2 [1,M,3] -> R [1,3,3], t [1,3,1] (direction pc1 -> pc2).  ``icp=False`` returns the
        pose before the ICP refinement the reference always applies."""
        n_in = self.cfg["shape_priors"]["n_input_point"]
        with torch.no_grad():
            _, pc1 = farthest_point_sample(pc1_full.transpose(1, 2), n_in)
            _, pc2 = farthest_point_sample(pc2_full.transpose(1, 2), n_in)
            code1 = self.model.encode(pc1)
            code2 = self.model.encode(pc2)
            R, t, _, _ = kabsch_transformation_estimation(code1["z_so3"] + code1["t"], code2["z_so3"] + code2["t"])
        if optim:
            from .optim import refine_registration

            R, t = refine_registration(self.model, pc1.transpose(1, 2), pc2.transpose(1, 2), code1, code2, R, t,
                                       self.cfg["registration"])
        if not icp:
            return R, t
        with torch.no_grad():
            # ICP refinement on the sub-sampled clouds, initialised with the pose so far (more_solver.py:182-189)
            s0 = torch.ones(R.shape[0], device=R.device)
            sol = iterative_closest_point(pc1.transpose(1, 2), pc2.transpose(1, 2),
                                          init_transform=SimilarityTransform(R.transpose(-1, -2), t.squeeze(2), s0))
            R, t, _ = sol.RTs
        return R.transpose(-1, -2), t.unsqueeze(2)

    def _optimize_code(self, code, pc, mask):
        """more_solver.py:191-228: pc [3,Nmax], mask [1,Nmax] -> the code with the lowest SDF loss."""
        from .optim import optimize_code

        n_in = self.cfg["shape_priors"]["n_input_point"]
        sub, _ = farthest_point_sample_masked(pc[None], mask.reshape(1, -1), n_in)
        return optimize_code(self.model, code, sub.transpose(1, 2))

    def _transform_latent(self, code, tsfm):
        """more_solver.py:230-244: rotate the equivariant code and move the centre by tsfm [b,3,4]."""
        R = tsfm[:, :, :3]
        return {"z_so3": (code["z_so3"] @ R.transpose(-1, -2)).detach().clone(),
                "z_inv": code["z_inv"].detach().clone(),
                "t": se3_transform(tsfm, code["t"]).detach().clone(),
                "s": code["s"].detach().clone()}

    # ------------------------------------------------------------------ meshes
    @property
    def mesh_extractor(self):
        if self._mesh_extractor is None:
            from .mesh_extractor import Generator3D

            self._mesh_extractor = Generator3D(**self.cfg["mesh_extractor"])
        return self._mesh_extractor

    def _mesh_from_latent(self, latent_code):
        """more_solver.py:37-58: extract in the canonical frame (t = 0, s = 1), then scale and translate.
        Returns (vertices [V,3], faces [F,3]) tensors on the device (the reference returns a trimesh)."""
        canon = dict(latent_code)
        canon["t"] = torch.zeros_like(latent_code["t"])
        canon["s"] = torch.ones_like(latent_code["s"])
        v, f = self.mesh_extractor.generate_from_latent(canon, self.model.decoder)
        v = v * latent_code["s"].reshape(1, 1) + latent_code["t"].reshape(1, 3)
        return v, f

    def _mesh_from_pc(self, pc):
        _, pc_down = farthest_point_sample(pc.transpose(1, 2), self.cfg["shape_priors"]["n_input_point"])
        return self._mesh_from_latent(self.model.encode(pc_down))

    # ------------------------------------------------------------------ end to end
    def _solve_end2end(self, ref, rescan, optim=False, with_mesh=False, icp=True):
        """more_solver.py:246-290.  ``ref`` / ``rescan``: {"pc": [n,3,Nmax], "pc_mask": [n,1,Nmax] bool} (tensors or
        lists of per-instance tensors of one common Nmax).  Returns the reference's dict: ``matches`` (matches0),
        ``registration`` (list of [1,3,4] SE(3) or None), ``ref_pc_lst`` / ``rescan_pc_lst`` (the FPS'ed clouds
        [n,3,n_in] instead of the reference's ragged full clouds) and, with ``with_mesh``, ``mesh_lst``."""
        if ref is None:
            return None
        n_in = self.cfg["shape_priors"]["n_input_point"]
        stack = lambda v: v if torch.is_tensor(v) else torch.stack(list(v), 0)
        with torch.no_grad():
            ref_sub, _ = farthest_point_sample_masked(stack(ref["pc"]), stack(ref["pc_mask"]), n_in)
            res_sub, _ = farthest_point_sample_masked(stack(rescan["pc"]), stack(rescan["pc_mask"]), n_in)
            ref_codes = self.model.encode(ref_sub)
            res_codes = self.model.encode(res_sub)
            matches = self._solve_object_matching(ref_codes, res_codes, "sequential")
            m0 = matches["matches0"]
            R, t, _ = kabsch_from_codes(ref_codes, res_codes, m0)           # all matched pairs, one launch
        out = {"ref_pc_lst": ref_sub, "rescan_pc_lst": res_sub, "matches": m0, "ref_codes": ref_codes,
               "rescan_codes": res_codes}
        ok = (m0 >= 0).nonzero().reshape(-1)
        reg = [None] * m0.shape[0]
        if ok.numel():
            mi = m0[ok]
            p1, p2 = ref_sub[ok].transpose(1, 2).contiguous(), res_sub[mi].transpose(1, 2).contiguous()
            Rk, tk = R[ok], t[ok]
            if optim:
                from .optim import refine_registration

                c1 = {k: v[ok] for k, v in ref_codes.items()}
                c2 = {k: v[mi] for k, v in res_codes.items()}
                Rk, tk = refine_registration(self.model, p1, p2, c1, c2, Rk, tk, self.cfg["registration"])
            if icp:
                with torch.no_grad():
                    s0 = torch.ones(ok.numel(), device=R.device)
                    sol = iterative_closest_point(p1, p2, init_transform=SimilarityTransform(Rk.transpose(-1, -2),
                                                                                             tk.squeeze(2), s0))
                    Rk, tk = sol.RTs.R.transpose(-1, -2), sol.RTs.T.unsqueeze(2)
            g = Rt_to_SE3(Rk, tk)
            for j, i in enumerate(ok.tolist()):
                reg[i] = g[j:j + 1]
        out["registration"] = reg
        if with_mesh:
            meshes = []
            for i, g in enumerate(reg):
                if g is None:
                    meshes.append(None)
                    continue
                j = int(m0[i])
                cur = {k: res_codes[k][j][None] for k in ("z_so3", "z_inv", "s", "t")}
                meshes.append(self._mesh_from_latent(self._transform_latent(cur, se3_inverse(g))))
            out["mesh_lst"] = meshes
        return out

    @torch.no_grad()
    def solve_scene_pair(self, ref_pcs, rescan_pcs, method="sequential"):
        """ref_pcs [n,3,N], rescan_pcs [m,3,N]: encode both sets (one batched call each), match on the
        invariant codes, fit one SE(3) per matched ref instance from the equivariant codes
        (the core of ``_solve_end2end`` for pre-sampled clouds)."""
        ref = self.model.encode(ref_pcs)
        res = self.model.encode(rescan_pcs)
        matches = self._solve_object_matching(ref, res, method)
        R, t, resid = kabsch_from_codes(ref, res, matches["matches0"])
        return {"ref_codes": ref, "rescan_codes": res, "matches": matches, "R": R, "t": t, "res": resid}
