"""Oracle: one paired-image training step and the traversal chains (test infrastructure, CPU torch).

Follows /root/reference/lib/trainer.py:184-261 (loop body) with every random draw *injected*
(z, target path indices, target shift magnitudes), because the reference's RNG is device-default and
unseeded; and /root/reference/traverse_latent_space.py:361-463 for the traversal.
"""
import torch
import torch.nn.functional as F

from . import support_sets as o_ss
from . import reconstructor as o_rec
from . import stylegan2 as o_sg2
from . import proggan as o_pg
from . import sngan as o_sn
from . import biggan as o_bg


def make_generator(gan_type, g_state, **kw):
    """Returns (G(z, shift) -> image, get_w or None) closures over a frozen generator state."""
    if gan_type == 'StyleGAN2':
        size = kw.get('size', 1024)
        wspace = kw.get('shift_in_w_space', False)
        gen = lambda z, shift=None, latent_is_w=False: o_sg2.generate(
            g_state, z, shift, size=size, shift_in_w_space=wspace, latent_is_w=latent_is_w)
        return gen, (lambda z: o_sg2.mapping(g_state, z))
    if gan_type == 'ProgGAN':
        plan = kw.get('plan', o_pg.PLAN_1024)
        return (lambda z, shift=None: o_pg.generate(g_state, z, shift, plan=plan)), None
    if gan_type == 'SNGAN':
        model = kw.get('model', 'sn_resnet32')
        return (lambda z, shift=None: o_sn.generate(g_state, z, shift, model=model)), None
    if gan_type == 'BigGAN':
        classes = kw['classes']
        res = kw.get('resolution', 128)
        ch = kw.get('ch', 96)
        return (lambda z, shift=None: o_bg.generate(g_state, z, classes, shift, resolution=res, ch=ch)), None
    raise ValueError(gan_type)


def sample_shift_magnitudes(batch, min_mag, max_mag, generator=None):
    """lib/trainer.py:212-221 including the index-weighted multinomial draw (Appendix B.1)."""
    pos = (min_mag - max_mag) * torch.rand(batch, generator=generator) + max_mag
    neg = (min_mag - max_mag) * torch.rand(batch, generator=generator) - min_mag
    pool = torch.cat((neg, pos))
    ids = torch.arange(len(pool), dtype=torch.float)
    return pool[torch.multinomial(ids, batch, replacement=False, generator=generator)]


def paired_step(gen, s_state, r_state, z, indices, magnitudes, *, reconstructor_type='ResNet',
                learn_gammas=True, lambda_cls=1.0, lambda_reg=0.25, get_w=None, running=None, train_bn=True):
    """Forward + backward of one step.  Returns dict(img, img_shifted, shift, logits, pred_mag,
    cls_loss, reg_loss, loss, accuracy, grads={'S': {...}, 'R': {...}}).  train_bn=False evaluates the Reconstructor's
    BatchNorm with its running statistics (R.eval()) - used by the well-conditioned whole-graph gradient test."""
    K = s_state['SUPPORT_SETS'].shape[0]
    s_leaf = {k: v.detach().clone().requires_grad_(k != 'ALPHAS') for k, v in s_state.items()}
    if not learn_gammas:
        s_leaf['LOGGAMMA'].requires_grad_(False)
    r_keys = o_rec.trainable_keys(r_state)
    r_leaf = dict(r_state)
    for k in r_keys:
        r_leaf[k] = r_state[k].detach().clone().requires_grad_(True)

    img = gen(z)                                                            # trainer.py:200
    mask = o_ss.one_hot(indices, K, z.dtype)                                # :227-231
    where = get_w(z) if get_w is not None else z                            # :236
    direction = o_ss.forward(s_leaf, mask, where, learn_gammas=learn_gammas)
    shift = magnitudes.reshape(-1, 1) * direction                           # :235
    img_shifted = gen(z, shift)                                             # :239
    logits, pred = o_rec.forward(r_leaf, img, img_shifted, reconstructor_type, train_bn, running)   # :242
    cls = F.cross_entropy(logits, indices)                                  # :245
    reg = torch.mean(torch.abs(pred - magnitudes))                          # :246
    loss = lambda_cls * cls + lambda_reg * reg                              # :249
    loss.backward()                                                         # :250
    acc = (logits.argmax(dim=1) == indices).float().mean()
    grads = {
        'S': {k: v.grad for k, v in s_leaf.items() if v.grad is not None},
        'R': {k: r_leaf[k].grad for k in r_keys if r_leaf[k].grad is not None},
    }
    return dict(img=img.detach(), img_shifted=img_shifted.detach(), shift=shift.detach(),
                logits=logits.detach(), pred_mag=pred.detach(), cls_loss=cls.detach(), reg_loss=reg.detach(),
                loss=loss.detach(), accuracy=acc, grads=grads)


def adam_update(param, grad, exp_avg, exp_avg_sq, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (lib/trainer.py:153,156): in-place update, returns nothing."""
    exp_avg.mul_(b1).add_(grad, alpha=1 - b1)
    exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (exp_avg_sq.sqrt() / (bc2 ** 0.5)).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-lr / bc1)


def traverse_chain(s_state, start, path, eps, shift_steps, learn_gammas=True, shift_leap=1):
    """One (latent, path) chain of traverse_latent_space.py:361-438.

    start: [1, d] (z, or w when shifting in W space).  Returns (codes [2*n+1, d], shifts [2*n+1, d])
    ordered from the most negative step to the most positive, centre = (start, 0)."""
    K = s_state['SUPPORT_SETS'].shape[0]
    mask = torch.zeros(1, K, dtype=start.dtype)
    mask[0, path] = 1.0
    codes, shifts = [start.clone()], [torch.zeros_like(start)]
    for sign in (1.0, -1.0):
        cur = start.clone()
        cnt = 0
        for _ in range(shift_steps):
            cnt += 1
            shift = sign * eps * o_ss.forward(s_state, mask, cur, learn_gammas=learn_gammas)
            cur = cur + shift
            if cnt == shift_leap:
                if sign > 0:
                    codes.append(cur)
                    shifts.append(shift)
                else:
                    codes.insert(0, cur)
                    shifts.insert(0, shift)
                cnt = 0
    return torch.cat(codes), torch.cat(shifts)
