"""More_Solver -- the inference orchestration of lib_more/more_solver.py on the CUDA hot path.

Built: ``_solve_object_matching`` (more_solver.py:71-93, methods "sequential" and "nn"),
``_solve_pairwise_registration(pc1, pc2, optim=False)`` (:95-116,182-189: FPS to n_input_point, encode both,
Kabsch on z_so3 + t, ICP refinement) and a batched ``solve_scene_pair`` that does encode -> match -> pose for two
instance sets without leaving the GPU.  Out of scope: the ``optim=True`` SE(3) Adam
loop, ``_optimize_code`` and mesh extraction.
"""
from __future__ import annotations

import torch

from .matcher_new import nn_matcher, sequential_matcher
from .ops import SimilarityTransform, farthest_point_sample, iterative_closest_point
from .pose_estimation import kabsch_from_codes, kabsch_transformation_estimation


class More_Solver:
    def __init__(self, model, cfg=None) -> None:
        """``model`` is a ``livingscenes_b200.Shape_Prior``; ``cfg`` follows configs/more_3rscan.yaml
        (only ``shape_priors.n_input_point`` and ``fps.n_init`` are read)."""
        self.model = model
        self.cfg = cfg or {"shape_priors": {"n_input_point": 1024}, "fps": {"n_init": 1, "random_start": False}}
        if self.cfg["fps"].get("n_init", 1) != 1:
            raise NotImplementedError("fps.n_init > 1 (random restarts) is not built")

    def _solve_object_matching(self, src_codes, tgt_codes, method):
        inv_src = src_codes["z_inv"].detach()
        inv_tgt = tgt_codes["z_inv"].detach()
        if method == "nn":
            return nn_matcher(inv_src.T[None], inv_tgt.T[None])
        if method == "sequential":
            return sequential_matcher(inv_src, inv_tgt)
        raise NotImplementedError(f"matching method {method!r} is not used by the evals and not built "
                                  "(sinkhorn / sim3_seq / eq_seq: SURVEY.md 8f rank 4)")

    @torch.no_grad()
    def _solve_pairwise_registration(self, pc1_full, pc2_full, optim=False, icp=True):
        """pc1 [1,N,3], pc2 [1,M,3] -> R [1,3,3], t [1,3,1] (direction pc1 -> pc2).  ``icp=False`` returns the
        code-based Kabsch pose without the ICP refinement the reference always applies."""
        if optim:
            raise NotImplementedError("optim=True (SE(3) Adam refinement) is out of scope (SURVEY.md 8f)")
        n_in = self.cfg["shape_priors"]["n_input_point"]
        _, pc1 = farthest_point_sample(pc1_full.transpose(1, 2), n_in)
        _, pc2 = farthest_point_sample(pc2_full.transpose(1, 2), n_in)
        code1 = self.model.encode(pc1)
        code2 = self.model.encode(pc2)
        R, t, _, _ = kabsch_transformation_estimation(code1["z_so3"] + code1["t"], code2["z_so3"] + code2["t"])
        if not icp:
            return R, t
        # ICP refinement on the sub-sampled clouds, initialised with the code-based pose (more_solver.py:182-189)
        s0 = torch.ones(R.shape[0], device=R.device)
        sol = iterative_closest_point(pc1.transpose(1, 2), pc2.transpose(1, 2),
                                      init_transform=SimilarityTransform(R.transpose(-1, -2), t.squeeze(2), s0))
        R, t, _ = sol.RTs
        return R.transpose(-1, -2), t.unsqueeze(2)

    @torch.no_grad()
    def solve_scene_pair(self, ref_pcs, rescan_pcs, method="sequential"):
        """ref_pcs [n,3,N], rescan_pcs [m,3,N]: encode both sets (one batched call each), match on the
        invariant codes, fit one SE(3) per matched ref instance from the equivariant codes
        (the core of ``_solve_end2end``, more_solver.py:246-282, without the mesh stage)."""
        ref = self.model.encode(ref_pcs)
        res = self.model.encode(rescan_pcs)
        matches = self._solve_object_matching(ref, res, method)
        R, t, resid = kabsch_from_codes(ref, res, matches["matches0"])
        return {"ref_codes": ref, "rescan_codes": res, "matches": matches, "R": R, "t": t, "res": resid}
