"""DeepSDF decoder + FieldWrapper -- drop-ins for the reference's SDF query path.

Mirrors ``lib_shape_prior/core/lib/implicit_func/deepsdf_decoder.py:9-123`` (parameter names
``lin{0..7}.{bias,weight_g,weight_v}``, ``lin8.{weight,bias}``) and ``model_utils.py:221-263``
(``FieldWrapper.forward(query, z_none, c, return_sdf=False)``, ``inner_deepsdf`` branch).
The modules hold parameters; the arithmetic runs in ``ls_sdf_decode`` (C ABI).
Back-propagation to the query points and the code (the reference's ``optim=True`` registration and
``_optimize_code``, more_solver.py:118-228, use autograd through the decoder) runs in ``ls_sdf_backward``
behind a ``torch.autograd.Function``; the decoder WEIGHTS get no gradient (inference-time optimisation only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import distributions as dist
from torch import nn

from . import _lib


class _WNLinear(nn.Module):
    """Parameters of ``nn.utils.weight_norm(nn.Linear(i, o))``: bias, weight_g [o,1], weight_v [o,i]."""

    def __init__(self, i: int, o: int):
        super().__init__()
        lin = nn.Linear(i, o)
        self.bias = nn.Parameter(lin.bias.detach().clone())
        self.weight_g = nn.Parameter(lin.weight.detach().norm(dim=1, keepdim=True))
        self.weight_v = nn.Parameter(lin.weight.detach().clone())

    def effective(self):
        return torch._weight_norm(self.weight_v.detach(), self.weight_g.detach(), 0), self.bias.detach()


class _Linear(nn.Module):
    def __init__(self, i: int, o: int):
        super().__init__()
        lin = nn.Linear(i, o)
        self.weight = nn.Parameter(lin.weight.detach().clone())
        self.bias = nn.Parameter(lin.bias.detach().clone())

    def effective(self):
        return self.weight.detach(), self.bias.detach()


class DeepSDF_Decoder(nn.Module):
    def __init__(self, latent_size, dims, dropout=None, dropout_prob=0.0, norm_layers=(), latent_in=(),
                 weight_norm=False, xyz_in_all=None, use_tanh=False, latent_dropout=False, pe_dim=3):
        super().__init__()
        if xyz_in_all or use_tanh or latent_dropout:
            raise NotImplementedError("only the shipped decoder configuration is built")
        if not weight_norm and len(list(norm_layers or ())) > 0:
            # the reference inserts nn.LayerNorm ("bn{l}") for these layers (deepsdf_decoder.py:59-63): not built
            raise NotImplementedError("norm_layers without weight_norm (LayerNorm variant) is not built")
        full = [latent_size + pe_dim] + list(dims) + [1]
        self.latent_size, self.pe_dim = latent_size, pe_dim
        self.num_layers = len(full)
        self.latent_in = list(latent_in)
        self.in_dims, self.out_dims = [], []
        for layer in range(self.num_layers - 1):
            out_dim = full[layer + 1] - full[0] if layer + 1 in self.latent_in else full[layer + 1]
            wn = weight_norm and layer in list(norm_layers)
            setattr(self, f"lin{layer}", (_WNLinear if wn else _Linear)(full[layer], out_dim))
            self.in_dims.append(full[layer])
            self.out_dims.append(out_dim)
        if self.num_layers - 1 != 9 or self.latent_in != [4] or pe_dim != latent_size + 1:
            raise NotImplementedError("ls_sdf_decode is built for the shipped 9-layer inner_deepsdf decoder "
                                      "(latent_in=[4], pe_dim=latent_size+1)")
        self._packed: Optional[dict] = None

    def _pack(self, device) -> dict:
        ver = (tuple((p.data_ptr(), p._version) for p in self.parameters()), str(device))
        if self._packed is not None and self._packed["ver"] == ver:
            return self._packed
        L, H = self.latent_size, self.out_dims[0]
        arrays = {}
        eff = [getattr(self, f"lin{l}").effective() for l in range(9)]
        W0, W4 = eff[0][0].float().cpu(), eff[4][0].float().cpu()
        h3 = self.out_dims[3]
        arrays["w0_zinv"] = W0[:, :L]
        arrays["w4_zinv"] = W4[:, h3:h3 + L]
        mats = {0: W0[:, L:], 4: torch.cat([W4[:, :h3], W4[:, h3 + L:]], 1)}
        for l in range(9):
            W = mats.get(l, eff[l][0].float().cpu())
            K = W.shape[1]
            Kp = (K + 7) // 8 * 8
            Wp = torch.zeros(W.shape[0], Kp)
            Wp[:, :K] = W
            arrays[f"w{l}"] = Wp
            arrays[f"b{l}"] = eff[l][1].float().cpu()
        offs, total = {}, 0
        for k, v in arrays.items():
            offs[k] = total
            total += (v.numel() + 63) // 64 * 64
        blob = torch.zeros(total, dtype=torch.float32)
        for k, v in arrays.items():
            blob[offs[k]:offs[k] + v.numel()] = v.reshape(-1)
        blob = blob.to(device)
        base = blob.data_ptr()
        d = _lib.DecoderDesc()
        d.latent, d.hidden, d.n_layers, d.latent_in = L, H, 9, 4
        for l in range(9):
            d.w[l] = base + 4 * offs[f"w{l}"]
            d.b[l] = base + 4 * offs[f"b{l}"]
            d.out_dims[l] = self.out_dims[l]
            d.in_dims[l] = mats[l].shape[1] if l in mats else self.in_dims[l]
        d.w0_zinv = base + 4 * offs["w0_zinv"]
        d.w4_zinv = base + 4 * offs["w4_zinv"]
        tc = {}
        for l in range(8):
            K = d.in_dims[l]
            Kp = (K + 7) // 8 * 8
            wv = blob[offs[f"w{l}"]:offs[f"w{l}"] + self.out_dims[l] * Kp].view(self.out_dims[l], Kp)[:, :K]
            tc[l] = _lib.tc_pack(wv.contiguous())
            d.w_tc[l] = tc[l].data_ptr()
        self._packed = {"ver": ver, "blob": blob, "desc": d, "tc": tc, "mats": {l: mats.get(l, eff[l][0].float().cpu())
                                                                                  for l in range(8)}, "bwd": False}
        return self._packed

    def _pack_backward(self, device) -> dict:
        """Transposed weights for ``ls_sdf_backward`` (built on first use: inference never pays for them)."""
        pk = self._pack(device)
        if pk["bwd"]:
            return pk
        d, h3 = pk["desc"], self.out_dims[3]
        keep = []

        def put(Wt):  # [in][out] -> device copy padded to a multiple of 8 columns + its tcgen05 image
            K = Wt.shape[1]
            Kp = (K + 7) // 8 * 8
            Wp = torch.zeros(Wt.shape[0], Kp)
            Wp[:, :K] = Wt
            Wp = Wp.to(device)
            packed = _lib.tc_pack(Wp[:, :K].contiguous())
            keep.extend([Wp, packed])
            return Wp.data_ptr(), packed.data_ptr()

        for l in (0, 1, 2, 3, 5, 6, 7):
            d.wt[l], d.wt_tc[l] = put(pk["mats"][l].t().contiguous())
        W4 = pk["mats"][4]  # [H][h3 + L + 1]
        d.wt4_h, d.wt4_h_tc = put(W4[:, :h3].t().contiguous())
        d.wt4_u, d.wt4_u_tc = put(W4[:, h3:].t().contiguous())
        pk["bwd_keep"], pk["bwd"] = keep, True
        return pk

    @torch.no_grad()
    def query(self, query: torch.Tensor, code: dict) -> torch.Tensor:
        """query [B,M,3] world coordinates + code dict -> sdf [B,M] (tanh output)."""
        _lib.require_cuda(query, "query")
        dev = query.device
        q = query.detach().float().contiguous()
        B, M0, _ = q.shape
        if M0 % 4:
            # the tcgen05 GEMMs need 16-byte aligned column groups; an odd point count would send the whole call through
            # the FP32 SIMT GEMM, whose results differ from the 3xTF32 ones in the last bits.  Padding keeps the value of a
            # point independent of how many points share its call (the MISE loop queries ragged lists).
            q = torch.cat([q, q[:, -1:].expand(B, 4 - M0 % 4, 3)], 1).contiguous()
        M = q.shape[1]
        z_so3 = code["z_so3"].detach().float().contiguous()
        z_inv = code["z_inv"].detach().float().contiguous()
        s = code["s"].detach().float().reshape(B).contiguous()
        t = code["t"].detach().float().reshape(B, 3).contiguous()
        assert z_so3.shape == (B, self.latent_size, 3) and z_inv.shape == (B, self.latent_size)
        with torch.cuda.device(dev):
            pk = self._pack(dev)
            nbytes = C.c_size_t(0)
            _lib.check(_lib.lib().ls_sdf_workspace_bytes(C.byref(pk["desc"]), B, M, C.byref(nbytes)),
                       "ls_sdf_workspace_bytes")
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            sdf = torch.empty(B, M, device=dev)
            rc = _lib.lib().ls_sdf_decode(C.byref(pk["desc"]), q.data_ptr(), z_so3.data_ptr(), z_inv.data_ptr(),
                                          s.data_ptr(), t.data_ptr(), B, M, sdf.data_ptr(), ws.data_ptr(),
                                          ws.numel(), _lib.stream_ptr(dev))
            _lib.check(rc, "ls_sdf_decode")
            _lib.launch_count += 1
            ws.record_stream(torch.cuda.current_stream(dev))
        return sdf if M == M0 else sdf[:, :M0].contiguous()

    BWD_MAX_COLS = 131072

    @torch.no_grad()
    def query_backward(self, query, code, grad_sdf, need=(True, True, True, True, True)):
        """Gradients of ``query`` (the SDF) for dLoss/dsdf = grad_sdf [B,M]: (grad_query [B,M,3], grad_z_so3, grad_z_inv,
        grad_s [B], grad_t [B,3]); long query lists are chunked and the code gradients summed."""
        _lib.require_cuda(query, "query")
        dev = query.device
        q = query.detach().float().contiguous()
        B, M, _ = q.shape
        z_so3 = code["z_so3"].detach().float().contiguous()
        z_inv = code["z_inv"].detach().float().contiguous()
        s = code["s"].detach().float().reshape(B).contiguous()
        t = code["t"].detach().float().reshape(B, 3).contiguous()
        g = grad_sdf.detach().float().contiguous()
        Mc = max(1, min(M, self.BWD_MAX_COLS // B))
        gq = torch.empty(B, M, 3, device=dev)
        acc = None
        with torch.cuda.device(dev):
            pk = self._pack_backward(dev)
            for m0 in range(0, M, Mc):
                m1 = min(M, m0 + Mc)
                qc, gc = q[:, m0:m1].contiguous(), g[:, m0:m1].contiguous()
                nbytes = C.c_size_t(0)
                _lib.check(_lib.lib().ls_sdf_backward_workspace_bytes(C.byref(pk["desc"]), B, m1 - m0, C.byref(nbytes)),
                           "ls_sdf_backward_workspace_bytes")
                ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
                gqc = torch.empty(B, m1 - m0, 3, device=dev)
                part = [torch.empty(B, self.latent_size, 3, device=dev), torch.empty(B, self.latent_size, device=dev),
                        torch.empty(B, device=dev), torch.empty(B, 3, device=dev)]
                rc = _lib.lib().ls_sdf_backward(C.byref(pk["desc"]), qc.data_ptr(), z_so3.data_ptr(), z_inv.data_ptr(),
                                                s.data_ptr(), t.data_ptr(), B, m1 - m0, gc.data_ptr(), gqc.data_ptr(),
                                                part[0].data_ptr(), part[1].data_ptr(), part[2].data_ptr(),
                                                part[3].data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
                _lib.check(rc, "ls_sdf_backward")
                _lib.launch_count += 1
                ws.record_stream(torch.cuda.current_stream(dev))
                gq[:, m0:m1] = gqc
                acc = part if acc is None else [a + b for a, b in zip(acc, part)]
        return (gq, *acc)

    def forward(self, input, phase="val"):
        raise NotImplementedError(
            "the CUDA decoder never materialises the [B,M,513] input tensor; call FieldWrapper.forward "
            "(query, None, code, return_sdf) as the reference's callers do (model_utils.py:230-263)")


class _SDFQuery(torch.autograd.Function):
    """sdf = decoder(query; z_so3, z_inv, s, t) with the CUDA forward (ls_sdf_decode) and backward (ls_sdf_backward)."""

    @staticmethod
    def forward(ctx, decoder, query, z_so3, z_inv, s, t):
        ctx.decoder = decoder
        ctx.save_for_backward(query, z_so3, z_inv, s, t)
        return decoder.query(query, {"z_so3": z_so3, "z_inv": z_inv, "s": s, "t": t})

    @staticmethod
    def backward(ctx, grad_sdf):
        query, z_so3, z_inv, s, t = ctx.saved_tensors
        gq, gzs, gzi, gs, gt = ctx.decoder.query_backward(query, {"z_so3": z_so3, "z_inv": z_inv, "s": s, "t": t}, grad_sdf)
        need = ctx.needs_input_grad
        return (None, gq.to(query.dtype) if need[1] else None, gzs.reshape(z_so3.shape) if need[2] else None,
                gzi.reshape(z_inv.shape) if need[3] else None, gs.reshape(s.shape) if need[4] else None,
                gt.reshape(t.shape) if need[5] else None)


class FieldWrapper(nn.Module):
    """model_utils.py:221-263 (inner_deepsdf branch)."""

    def __init__(self, decoder, decoder_type="inner_deepsdf", sdf2occ_factor=-1.0) -> None:
        super().__init__()
        if decoder_type != "inner_deepsdf":
            raise NotImplementedError("only decoder_type == 'inner_deepsdf' (the shipped model) is built")
        self.F = decoder
        self.sdf2occ_factor = sdf2occ_factor
        self.decoder_type = decoder_type

    def forward(self, query, z_none, c, return_sdf=False):
        if torch.is_grad_enabled() and any(torch.is_tensor(v) and v.requires_grad
                                           for v in (query, c["z_so3"], c["z_inv"], c["s"], c["t"])):
            # the reference's optimisation loops differentiate the SDF w.r.t. the query points / the code
            sdf = _SDFQuery.apply(self.F, query, c["z_so3"], c["z_inv"], c["s"], c["t"]).to(query.dtype)
        else:
            sdf = self.F.query(query, c).to(query.dtype)
        if return_sdf:
            return sdf
        return dist.Bernoulli(logits=self.sdf2occ_factor * sdf)
