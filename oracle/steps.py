"""ORACLE (test infrastructure, never imported by the product path) — train_step restatements.

One simultaneous generator + discriminator update as in the reference's four `train_step`s,
using torch autograd on the CPU restatements in `models.py` and the hand-written Keras Adam in
`ops_torch.py`.  `vgg` may be None: then content_loss is 0 (the "G+D step" of BASELINE.json's
metric); with seeded synthetic VGG19 weights it is the literal step (ImageNet weights cannot be
downloaded here).

PARITY UNPINNED.
"""
from __future__ import annotations

import torch

from . import models as M
from . import ops_torch as T


def _trainable(p):
    return {k: v for k, v in p.items() if not k.endswith(("moving_mean", "moving_variance"))}


def _grads(loss, params):
    names = list(params.keys())
    gs = torch.autograd.grad(loss, [params[k] for k in names], retain_graph=True, allow_unused=True)
    return dict(zip(names, gs))


def _prep(p):
    for k, v in p.items():
        v.requires_grad_(not k.endswith(("moving_mean", "moving_variance")))


def _finish(g, d, g_state, d_state, gen_grads, disc_grads, gen_opt, disc_opt):
    for p in (g, d):
        for v in p.values():
            v.requires_grad_(False)
    gen_opt.apply(g, gen_grads)
    disc_opt.apply(d, disc_grads)
    with torch.no_grad():
        for k, v in g_state.items():
            g[k].copy_(v)
        for k, v in d_state.items():
            d[k].copy_(v)


def srgan_train_step(g, d, vgg, gen_opt, disc_opt, x, y, *, fsrgan=False, acts=None, out=None, q=None):
    """train_srgan.py:61-118 (fsrgan=False) and train_fsrgan.py:61-120 (fsrgan=True).
    D's BN moving statistics are updated twice (real call, then fake call)."""
    _prep(g); _prep(d)
    g_state, d_state = {}, {}
    gen = M.fsrgan_generator if fsrgan else M.srgan_generator
    qkw = {} if (q is None or fsrgan) else {"q": q}
    gen_output = gen(g, x, True, g_state, acts, **qkw)                               # :75
    disc_real = M.patch_discriminator(d, y, True, d_state, q=q)                      # :78
    disc_fake = M.patch_discriminator(d, gen_output, True, d_state, acts, q=q)       # :79 (acts: the FAKE call's layers)
    zero = torch.zeros((), dtype=x.dtype)
    content = M.content_loss(vgg, y, gen_output) if vgg is not None else zero        # :86
    adv = 1e-3 * T.bce_from_logits(disc_fake, 1.0)                                   # :87
    mse = T.mse(y, gen_output)                                                       # :88
    mae = T.mae(y, gen_output)                                                       # :89
    var = 1e-5 * T.total_variation_mean(y - gen_output)                              # :90
    gen_loss = content + adv + 0 * mse + mae + (0 * var if not fsrgan else 0)        # :91 / fsrgan :91
    disc_loss = T.bce_from_logits(disc_real, 1.0) + T.bce_from_logits(disc_fake, 0.0)  # :94-96
    if fsrgan:
        disc_loss = 0.5 * disc_loss                                                  # train_fsrgan.py:96
    gen_grads = _grads(gen_loss, _trainable(g))                                      # :111
    disc_grads = _grads(disc_loss, _trainable(d))                                    # :112
    if out is not None:
        out.update(gen_output=gen_output.detach(), disc_real=disc_real.detach(), disc_fake=disc_fake.detach(),
                   gen_grads=gen_grads, disc_grads=disc_grads)
        if acts:
            names = [k for k, v in acts.items() if v.requires_grad]
            gs = torch.autograd.grad(gen_loss, [acts[k] for k in names], retain_graph=True, allow_unused=True)
            out["act_grads"] = {k: v for k, v in zip(names, gs) if v is not None}
    _finish(g, d, g_state, d_state, gen_grads, disc_grads, gen_opt, disc_opt)        # :115-116
    vals = [v.detach() for v in (gen_loss, adv, mae, mse, content, disc_loss, var)]
    gl, adv_, mae_, mse_, content_, dl, var_ = vals
    if fsrgan:
        return gl, gl, dl, adv_, content_, mse_, mae_, var_                          # train_fsrgan.py:120
    return gl, adv_, mae_, mse_, content_, dl, var_                                  # :118


def autoencoder_train_step(g, d, vgg, gen_opt, disc_opt, x, y, *, acts=None, out=None):
    """train_autoencoder.py:66-112; discriminator ends in sigmoid, BCE on probabilities."""
    _prep(g); _prep(d)
    g_state, d_state = {}, {}
    fake_hr = M.autoencoder_generator(g, x, True, g_state, acts)
    valid_pred = M.patch_discriminator(d, y, True, d_state, sigmoid=True)
    fake_pred = M.patch_discriminator(d, fake_hr, True, d_state, sigmoid=True)
    zero = torch.zeros((), dtype=x.dtype)
    content = M.content_loss(vgg, y, fake_hr) if vgg is not None else zero           # :90
    adv = 1e-3 * T.bce_from_probs(fake_pred, 1.0)                                    # :92
    mse = T.mse(y, fake_hr)
    mae = T.mae(y, fake_hr)
    perceptual = content + adv + 0 * mse + mae                                       # :95
    disc_loss = T.bce_from_probs(valid_pred, 1.0) + T.bce_from_probs(fake_pred, 0.0)  # :100-102
    gen_grads = _grads(perceptual, _trainable(g))
    disc_grads = _grads(disc_loss, _trainable(d))
    if out is not None:
        out.update(gen_output=fake_hr.detach(), disc_real=valid_pred.detach(), disc_fake=fake_pred.detach(),
                   gen_grads=gen_grads, disc_grads=disc_grads)
    _finish(g, d, g_state, d_state, gen_grads, disc_grads, gen_opt, disc_opt)
    return tuple(v.detach() for v in (disc_loss, adv, content, mse, mae))            # :112


def pix2pix_train_step(g, d, vgg, gen_opt, disc_opt, x, y, masks_main, masks_identity, *, acts=None, out=None):
    """train_pix2pix.py:33-71 with Pix2Pix.generator_loss (pix2pix.py:74-94) and discriminator_loss (:96-103).
    The identity term runs G a second time on the target in training mode (BN statistics and moving
    averages update again, dropout draws again)."""
    _prep(g); _prep(d)
    g_state, d_state = {}, {}
    gen_output = M.pix2pix_generator(g, x, True, g_state, acts, masks_main)           # :44
    disc_real = M.pix2pix_discriminator(d, x, y, True, d_state)                       # :47
    disc_fake = M.pix2pix_discriminator(d, x, gen_output, True, d_state)              # :48
    zero = torch.zeros((), dtype=x.dtype)
    gan = 1e-3 * T.bce_from_logits(disc_fake, 1.0)                                    # pix2pix.py:75
    var = 1e-5 * T.total_variation_mean(y - gen_output)                               # :78
    l1 = T.mae(y, gen_output)                                                         # :81
    l2 = T.mse(y, gen_output)                                                         # :84
    cont = M.content_loss(vgg, gen_output, y) if vgg is not None else zero            # :87
    ident = T.mae(M.pix2pix_generator(g, y, True, g_state, None, masks_identity), y)  # :90
    total = gan + l2 + cont + var + l1 + ident                                        # :92
    disc_loss = T.bce_from_logits(disc_real, 1.0) + T.bce_from_logits(disc_fake, 0.0)
    gen_grads = _grads(total, _trainable(g))
    disc_grads = _grads(disc_loss, _trainable(d))
    if out is not None:
        out.update(gen_output=gen_output.detach(), disc_real=disc_real.detach(), disc_fake=disc_fake.detach(),
                   gen_grads=gen_grads, disc_grads=disc_grads)
    _finish(g, d, g_state, d_state, gen_grads, disc_grads, gen_opt, disc_opt)
    return tuple(v.detach() for v in (total, gan, l1, l2, cont, disc_loss, var, ident))  # train_pix2pix.py:71
