"""Drop-in for the reference's `train_pix2pix.train_step` (train_pix2pix.py:33-71) with
Pix2Pix.generator_loss (pix2pix.py:74-94) and discriminator_loss (pix2pix.py:96-103)."""
from __future__ import annotations

import torch


def train_step(model, x, y):
    """x (degraded) and y (clean): [B,256,256,3] float32 NHWC CUDA tensors in [-1,1].
    total = 1e-3*BCE(1, D(x,G(x))) + mse + content + 1e-5*mean(TV(y-G(x))) + mae + mean|G(y)-y|   (pix2pix.py:92);
    the identity term runs the generator a second time in training mode (BN statistics, moving averages and
    dropout all run again, pix2pix.py:90).  Returns (gen_total_loss, gen_gan_loss, gen_l1_loss, gen_l2_loss,
    content_loss, disc_loss, var_loss, identity_loss), the order of train_pix2pix.py:71."""
    E = model.engine
    E.new_step()
    xv, yv = E.input(x), E.input(y)
    # Independent sub-graphs run on the engine's side stream: D(real) next to the first generator pass, and the identity
    # generator pass next to D(fake) and the image losses.  Joins keep the reference's order of the BatchNorm
    # moving-statistics updates (D: real then fake; G: pass 0 then the identity pass).
    with E.branch():
        disc_real = model.discriminator([xv, yv], training=True)                 # train_pix2pix.py:47
    gen_output = model.generator(xv, training=True, pass_id=0)                   # :44
    E.join()
    with E.branch():
        ident_out = model.generator(yv, training=True, pass_id=1)                # pix2pix.py:90
    disc_fake = model.discriminator([xv, gen_output], training=True)             # :48

    seeds_g = []
    if model.use_vgg:
        content, dgf, gf = model.content_loss(yv, gen_output)                    # pix2pix.py:87 (symmetric in its arguments)
        seeds_g.append((gf, dgf))
        content = content[0]
    else:
        content = torch.zeros((), dtype=torch.float32, device=E.device)
    gan_raw, g_adv = E.bce(disc_fake, 1.0, True, 1e-3, key="adv")                # pix2pix.py:75
    out3, dgen = E.image_losses(gen_output, y, 1.0, 1.0, 1e-5, key="img")         # :78-84 (mae, mse, 1e-5*TV)
    E.join()
    id3, dident = E.image_losses(ident_out, y, 1.0, 0.0, 0.0, key="ident")
    real_loss, g_real = E.bce(disc_real, 1.0, True, 1.0, key="dreal")             # :97
    fake_loss, g_fake = E.bce(disc_fake, 0.0, True, 1.0, key="dfake")             # :99
    seeds_g += [(disc_fake, g_adv), (gen_output, dgen), (ident_out, dident)]

    E.backward([(disc_real, g_real), (disc_fake, g_fake)], "d")                  # train_pix2pix.py:63
    comm, hook = model.comm, None
    if comm is not None:
        side = [E._side_stream]
        comm.begin(model.disc_params)
        comm.finish(model.disc_params, side)        # 11 MB: overlaps the generator backward pass
        comm.begin(model.gen_params)
        hook = lambda: comm.poll(model.gen_params, E.complete, side)    # 218 MB in ~9 reverse-layer-order buckets
    E.backward(seeds_g, "g", collect=E.grad_record, on_node=hook)                # :62
    scale = 1.0
    if comm is not None:
        comm.finish(model.gen_params, side)
        scale = 1.0 / model.world_size
        comm.wait_for(model.disc_params)
        model.disc_optimizer.apply(E, model.disc_params, scale)                  # :67
        comm.wait_for(model.gen_params)
        model.gen_optimizer.apply(E, model.gen_params, scale)                    # :66
    else:
        model.gen_optimizer.apply(E, model.gen_params, scale)                    # :66
        model.disc_optimizer.apply(E, model.disc_params, scale)                  # :67
    model.iterations += 1

    gan = 1e-3 * gan_raw[0]
    l1, l2, var = out3[0], out3[1], 1e-5 * out3[2]
    ident = id3[0]
    total = gan + l2 + content + var + l1 + ident
    disc_loss = real_loss[0] + fake_loss[0]
    model.last = dict(gen_output=gen_output, disc_real=disc_real, disc_fake=disc_fake, ident_out=ident_out)
    return total, gan, l1, l2, content, disc_loss, var, ident
