"""The G+D update shared by the SRGAN / Fast-SRGAN / autoencoder train steps.

Restates train_srgan.py:61-118: one forward of G, D(real), D(fake) [and VGG19 on both images when the
content loss is enabled], the loss terms, then the two `tape.gradient` calls (discriminator variables
from disc_loss, generator variables from gen_loss through D(fake) and VGG) computed from the SAME
forward, an optional data-parallel gradient all-reduce, and both Adam updates.
"""
from __future__ import annotations

import torch


def gan_step(model, img_input, img_target, *, from_logits: bool, disc_scale: float, w_mae=1.0, w_mse=0.0, w_tv=0.0):
    E = model.engine
    E.new_step()
    x = E.input(img_input)
    y = E.input(img_target)
    # D(real) does not depend on the generator: it runs on the side stream next to the generator forward and is joined
    # before D(fake), which keeps the reference's order of the two BatchNorm moving-statistics updates (real, then fake)
    with E.branch():
        disc_real = model.discriminator(y, training=True)              # train_srgan.py:78
    gen_output = model.generator(x, training=True)                     # :75
    E.join()
    disc_fake = model.discriminator(gen_output, training=True)         # :79

    seeds_g = []
    if model.use_vgg:
        content, dgf, gf = model.content_loss(y, gen_output)          # :86
        seeds_g.append((gf, dgf))
    else:
        content = None
    adv_raw, g_adv = E.bce(disc_fake, 1.0, from_logits, 1e-3, key="adv")          # :87
    out3, dgen = E.image_losses(gen_output, y.t, w_mae, w_mse, w_tv)              # :88-90
    real_loss, g_real = E.bce(disc_real, 1.0, from_logits, disc_scale, key="dreal")   # :94
    fake_loss, g_fake = E.bce(disc_fake, 0.0, from_logits, disc_scale, key="dfake")   # :95
    seeds_g += [(disc_fake, g_adv), (gen_output, dgen)]

    E.backward([(disc_real, g_real), (disc_fake, g_fake)], "d")        # :112
    comm, hook = model.comm, None
    if comm is not None:
        side = [E._side_stream]
        comm.begin(model.disc_params)
        comm.finish(model.disc_params, side)        # the discriminator arena overlaps the generator backward pass below
        comm.begin(model.gen_params)
        hook = lambda: comm.poll(model.gen_params, E.complete, side)    # reverse-layer-order buckets, launched as they complete
    E.backward(seeds_g, "g", collect=E.grad_record, on_node=hook)                              # :111

    scale = 1.0
    if comm is not None:
        comm.finish(model.gen_params, side)
        scale = 1.0 / model.world_size
        # the two updates are independent: the discriminator's runs while the generator's last bucket is still in flight
        comm.wait_for(model.disc_params)
        model.disc_optimizer.apply(E, model.disc_params, scale)         # :116
        comm.wait_for(model.gen_params)
        model.gen_optimizer.apply(E, model.gen_params, scale)           # :115
    else:
        model.gen_optimizer.apply(E, model.gen_params, scale)           # :115
        model.disc_optimizer.apply(E, model.disc_params, scale)         # :116
    model.iterations += 1

    # the returned scalars (train_srgan.py:86-99, 118) in ONE launch; the dict holds 0-d views of one 7-vector, which the captured step
    # hands out as a single device->host copy (graph.GraphedStep.packed)
    from . import _lib
    t = E.buf(("loss", "terms"), (7,), torch.float32)
    _lib.check(E.lib.dg_gan_loss_terms(E.ctx, _lib.ptr(content), adv_raw.data_ptr(), out3.data_ptr(), real_loss.data_ptr(), fake_loss.data_ptr(),
                                       float(w_mae), float(w_mse), float(w_tv / 1e-5 if w_tv else 0.0), float(disc_scale), t.data_ptr(), E.st))
    return dict(gen_loss=t[0], adv_loss=t[1], mae_loss=t[2], mse_loss=t[3], content_loss=t[4], disc_loss=t[5],
                var_loss=t[6], gen_output=gen_output, disc_real=disc_real, disc_fake=disc_fake)
