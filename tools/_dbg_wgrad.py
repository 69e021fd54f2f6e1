import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import _lib as L
lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
N, H, W, cin, cout, k = [int(v) for v in sys.argv[1:7]]
has_bias = int(sys.argv[7]) if len(sys.argv) > 7 else 1
x = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
dy = torch.randn(N, H, W, cout, device="cuda").to(torch.bfloat16)
cp = L.DgConvParams(k, k, 1, k // 2, k // 2, 0, 0.0)
tx, ty = L.tensor(x), L.tensor(dy)
nb = lib.dg_umma_conv2d_wgrad_workspace_bytes(C.byref(tx), C.byref(ty), C.byref(cp))
print("workspace", nb, flush=True)
big = torch.full((nb // 4 + 2 * (1 << 20),), 7.0, device="cuda")
ws = big[(1 << 20):(1 << 20) + nb // 4]
out = torch.full((k * k * cin * cout + 2 * 65536,), 7.0, device="cuda")
dw = out[65536:65536 + k * k * cin * cout]
db = torch.full((cout + 2048,), 7.0, device="cuda")
L.check(lib.dg_umma_conv2d_wgrad(ctx, C.byref(tx), C.byref(ty), dw.data_ptr(), db[1024:].data_ptr() if has_bias else None, C.byref(cp), 0, ws.data_ptr(), nb, st))
torch.cuda.synchronize()
print("ws guards intact:", bool((big[:1 << 20] == 7).all()), bool((big[(1 << 20) + nb // 4:] == 7).all()))
print("dw guards intact:", bool((out[:65536] == 7).all()), bool((out[65536 + k * k * cin * cout:] == 7).all()))
print("db guards intact:", bool((db[:1024] == 7).all()), bool((db[1024 + cout:] == 7).all()))
ref = torch.einsum("nhwc,nhwo->co", x.float()[:, :, :, :], dy.float())  # centre tap only
centre = dw.view(k, k, cin, cout)[k // 2, k // 2]
print("centre tap relerr", ((centre - ref).abs().max() / ref.abs().max()).item())
