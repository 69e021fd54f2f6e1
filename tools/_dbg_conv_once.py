import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from denoise_gan_b200 import _lib as L
lib = L.load(); ctx = L.ctx(0); st = L.stream_ptr()
N, H, W, cin, cout, k = [int(v) for v in (sys.argv[1:7] if len(sys.argv) > 6 else (16, 96, 96, 64, 64, 3))]
x = torch.randn(N, H, W, cin, device="cuda").to(torch.bfloat16)
y = torch.empty(N, H, W, cout, device="cuda", dtype=torch.bfloat16)
w = torch.randn(k, k, cin, cout, device="cuda") * 0.05
pk = torch.empty(w.numel(), dtype=torch.bfloat16, device="cuda")
L.check(lib.dg_umma_pack_weights(ctx, w.data_ptr(), pk.data_ptr(), k, k, cin, cout, 0, st))
cp = L.DgConvParams(k, k, 1, k // 2, k // 2, 0, 0.0)
tx, ty = L.tensor(x), L.tensor(y)
L.check(lib.dg_umma_conv2d_fwd(ctx, C.byref(tx), pk.data_ptr(), None, C.byref(ty), C.byref(cp), None, st))
torch.cuda.synchronize()
ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float().permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
print("max err", (y.float() - ref).abs().max().item(), "ref max", ref.abs().max().item())
