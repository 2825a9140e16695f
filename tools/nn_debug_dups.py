import torch, numpy as np, sys, ctypes
sys.path.insert(0, '.')
from rfnet_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
b, n, m = 4, 2048, 16384
x1 = torch.rand(b, n, 3, device='cuda') - 0.5
x2 = torch.rand(b, m, 3, device='cuda') - 0.5
src = torch.randint(0, 300, (b, n - 300), device='cuda')
x1[:, 300:] = torch.gather(x1[:, :300], 1, src[..., None].expand(-1, -1, 3))
wsb = lib.rfnet_nn_distance_workspace_bytes(b, n, m)
ws = torch.zeros(wsb, dtype=torch.uint8, device='cuda')
d1 = torch.empty(b, n, device='cuda'); i1 = torch.empty(b, n, dtype=torch.int32, device='cuda')
d2 = torch.empty(b, m, device='cuda'); i2 = torch.empty(b, m, dtype=torch.int32, device='cuda')
sc = torch.zeros(1, dtype=torch.int64, device='cuda')
p = lambda t: ctypes.c_void_p(t.data_ptr())
rc = lib.rfnet_nn_distance_stats(b, n, p(x1), m, p(x2), p(d1), p(i1), p(d2), p(i2), p(ws), wsb, 0, p(sc), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print('rc', rc, 'scans', sc.item())
keys = (8 * b * (n + m) + 15) // 16 * 16
pad = lambda v: (v + 15) // 16 * 16
f = ws[keys:].view(torch.float32).view(-1, 4)
cv0 = f[: b * pad(m)].view(b, pad(m), 4)
cv1 = f[b * pad(m): b * pad(m) + b * pad(n)].view(b, pad(n), 4)
meta = f[b * pad(m) + b * pad(n):].view(2, b, 4)
print('meta0', meta[0].tolist()); print('meta1', meta[1].tolist())
print('dir0 rows with inf norm per cloud', torch.isinf(cv0[..., 3]).sum(1).tolist())
print('dir1 rows with inf norm per cloud', torch.isinf(cv1[..., 3]).sum(1).tolist(), 'of', pad(n))
print('dir1 first 304 alive?', (~torch.isinf(cv1[0, :304, 3])).sum().item(), ' alive beyond 300:', (~torch.isinf(cv1[0, 300:, 3])).sum().item())
