"""Diagnostic for the tcgen05 GEMM: selector weights x index-coded activations expose layout permutations."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from livingscenes_b200 import _lib
dev = torch.device("cuda:0")
R, K, n, B = 128, 16, 128, 1
W = torch.zeros(R, K)
for r in range(R):
    W[r, r % K] = 1.0
X = (torch.arange(K)[:, None] * 1000.0 + torch.arange(n)[None, :]).reshape(1, K, n).contiguous()
Wd, Xd = W.to(dev), X.to(dev)
out = torch.full((B, R, n), -7.0, device=dev)
packed = _lib.tc_pack(Wd)
rc = _lib.lib().ls_vn_linear(Wd.data_ptr(), packed.data_ptr(), Xd.data_ptr(), out.data_ptr(), R, K, K, B, n, _lib.stream_ptr(dev))
torch.cuda.synchronize()
o = out.cpu()[0]
ref = W @ X[0]
print("rc", rc, "zeros?", float(o.abs().max()), "untouched(-7)?", int((o == -7).sum()))
print("row0 cols0-9 ", o[0, :10].tolist())
print("ref          ", ref[0, :10].tolist())
print("row1 cols0-9 ", o[1, :10].tolist())
print("row5 cols0-9 ", o[5, :10].tolist())
print("col0 rows0-19", o[:20, 0].tolist())
print("refcol0      ", ref[:20, 0].tolist())
print("max abs err", float((o - ref).abs().max()))
# second probe: W = all ones in column k0 only -> out[r][n] = X[k0][n]
for k0 in (0, 3, 4, 9):
    W2 = torch.zeros(R, K); W2[:, k0] = 1.0
    W2d = W2.to(dev); p2 = _lib.tc_pack(W2d)
    _lib.lib().ls_vn_linear(W2d.data_ptr(), p2.data_ptr(), Xd.data_ptr(), out.data_ptr(), R, K, K, B, n, _lib.stream_ptr(dev))
    torch.cuda.synchronize()
    print(f"k0={k0}: row0", out[0, 0, :6].tolist(), " row77", out[0, 77, :6].tolist(), " expect", X[0, k0, :6].tolist())
# third: X = ones -> out[r][n] = sum_k W[r][k] ; W[r][k] = r + 0.01*k
W3 = torch.arange(R)[:, None] * 1.0 + 0.01 * torch.arange(K)[None, :]
W3d = W3.to(dev); p3 = _lib.tc_pack(W3d)
X3 = torch.ones(1, K, n, device=dev)
_lib.lib().ls_vn_linear(W3d.data_ptr(), p3.data_ptr(), X3.data_ptr(), out.data_ptr(), R, K, K, B, n, _lib.stream_ptr(dev))
torch.cuda.synchronize()
print("rowsum probe: out[:10,0]", out[0, :10, 0].tolist(), " expect", W3.sum(1)[:10].tolist())
