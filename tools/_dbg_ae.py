import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
from denoise_gan_b200.autoencoder import Autoencoder
from denoise_gan_b200.train_autoencoder import train_step
from denoise_gan_b200.dataloader import synthetic_pair
from denoise_gan_b200 import engine as EG
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
orig = EG.Engine._wgrad
def patched(self, x, dy, w, b, lin, flops=0.0):
    try:
        r = orig(self, x, dy, w, b, lin, flops)
        torch.cuda.synchronize()
        return r
    except Exception as e:
        print("FAILED wgrad", w.name, tuple(x.shape), x.dtype, tuple(dy.shape), dy.dtype, lin.kh, lin.stride, "bias", b is not None, flush=True)
        raise
EG.Engine._wgrad = patched
m = Autoencoder(SimpleNamespace(crop_size=256, scale=1, lr=1e-3, fp16=1, vgg=0, seed=0, retrain=0))
x, y = synthetic_pair(B, 256, 1, step=0)
train_step(m, x.cuda(), y.cuda())
torch.cuda.synchronize()
print("ok")
