"""Oracle: one fastai training step around the reference Transformer-XL.  TEST INFRASTRUCTURE (plain PyTorch fp32, CPU).

Restates what ``Learner.fit`` does per batch for ``music_model_learner`` (``deep_music_genre.py:1784-1807``; fastai 1.0.61
``RNNLearner`` / ``RNNTrainer`` / ``CrossEntropyFlat`` / ``OptimWrapper`` are un-vendored):

    model.train(); out = model(x)                                  # deep_music_genre.py:1617-1647
    loss = CrossEntropyFlat(out[0], y)                             # mean over b*T
    loss += alpha * out[2][-1].float().pow(2).mean()               # AR  (RNNTrainer.on_backward_begin)
    h = out[1][-1]; loss += beta * (h[:,1:] - h[:,:-1]).float().pow(2).mean()   # TAR (h is detached memory: value only)
    loss.backward(); clip_grad_norm_; p.mul_(1 - wd*lr); Adam(betas=(0.9,0.99)).step()

Dropout: the product draws counter-based masks on the device; ``install_dropout_masks`` replaces every dropout module of
the oracle model by a multiplication with a caller-provided mask so that both sides see the same masks.
Known deviation (shared by the product): ``clip_grad_norm_`` here runs ONCE over all parameters (one global
norm).  fastai's ``MixedPrecision(clip=0.5)`` - the reference's only clip setting, notebook cell 62 - clips per master-parameter layer
group of ``tfmerXL_lm_split``, so the clipped update differs from the reference whenever a single group's norm exceeds the threshold
while the others do not.
Parity status: unpinned at the fastai boundary (the reference stores no training artefacts); this file is the definition
the CUDA path is tested against.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class FixedMask(nn.Module):
    "Stands in for nn.Dropout / RNNDropout: multiplies by a fixed mask (already scaled by 1/(1-p)); identity if mask is None."
    def __init__(self):
        super().__init__()
        self.mask = None

    def forward(self, x):
        if self.mask is None:
            return x
        return x * self.mask.to(x.dtype)


def install_dropout_masks(model):
    """Swap every dropout of oracle.txl's SequentialRNN for a FixedMask; returns a dict site -> list of modules:
    'embed' [1], 'attn' [L], 'res1' [L], 'ff' [L], 'res2' [L], 'out' [1]  (mask shapes: [b,T,d], [b,H,T,S], [b,T,d],
    [b,T,d_inner], [b,T,d], [b,1,d])."""
    enc, dec = model[0], model[1]
    sites = {'embed': [], 'attn': [], 'res1': [], 'ff': [], 'res2': [], 'out': []}
    enc.drop_emb = FixedMask(); sites['embed'].append(enc.drop_emb)
    for layer in enc.layers:
        layer.mhra.drop_att = FixedMask(); sites['attn'].append(layer.mhra.drop_att)
        layer.mhra.drop_res = FixedMask(); sites['res1'].append(layer.mhra.drop_res)
        ff = layer.ff.layers
        assert isinstance(ff[2], (nn.Dropout, FixedMask)) and isinstance(ff[4], (nn.Dropout, FixedMask))
        ff[2] = FixedMask(); sites['ff'].append(ff[2])
        ff[4] = FixedMask(); sites['res2'].append(ff[4])
    dec.output_dp = FixedMask(); sites['out'].append(dec.output_dp)
    return sites


def rnn_trainer_loss(model_out, y, alpha=2., beta=1.):
    "CrossEntropyFlat + RNNTrainer AR/TAR; returns (total, ce, ar, tar)."
    decoded, raw_outputs, outputs = model_out
    ce = F.cross_entropy(decoded.reshape(-1, decoded.size(-1)).float(), y.reshape(-1))
    ar = alpha * outputs[-1].float().pow(2).mean() if alpha != 0. else decoded.new_zeros(())
    tar = decoded.new_zeros(())
    h = raw_outputs[-1]
    if beta != 0. and h.dim() == 3 and h.size(1) > 1:
        tar = beta * (h[:, 1:] - h[:, :-1]).float().pow(2).mean()
    return ce + ar + tar, ce, ar, tar


class AdamTrueWD:
    "fastai OptimWrapper(Adam, true_wd=True): p *= (1 - lr*wd), then torch.optim.Adam's update (bias-corrected, eps outside sqrt)."
    def __init__(self, params, betas=(0.9, 0.99), eps=1e-8):
        self.params = [p for p in params]
        self.betas, self.eps = betas, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    @torch.no_grad()
    def step(self, lr, wd=0.01, clip=None, betas=None):
        b1, b2 = betas or self.betas
        self.t += 1
        gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in self.params if p.grad is not None))
        coef = 1.0
        if clip:
            coef = min(1.0, clip / (gn.item() + 1e-6))
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad * coef
            p.mul_(1 - lr * wd)
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            mh, vh = m / (1 - b1 ** self.t), v / (1 - b2 ** self.t)
            p.addcdiv_(mh, vh.sqrt() + self.eps, value=-lr)
        return gn.item()


def unique_params(model):
    seen, out = set(), []
    for p in model.parameters():
        if id(p) not in seen:
            seen.add(id(p)); out.append(p)
    return out


def train_step(model, x, y, opt, lr, wd=0.01, clip=0.5, alpha=2., beta=1., betas=None):
    "One reference training step (model must be in train() mode with FixedMask dropouts or p=0); returns the loss parts."
    for p in unique_params(model):
        p.grad = None
    out = model(x)
    total, ce, ar, tar = rnn_trainer_loss(out, y, alpha, beta)
    total.backward()
    gn = opt.step(lr, wd=wd, clip=clip, betas=betas)
    return {'loss': total.item(), 'ce': ce.item(), 'ar': float(ar.detach()), 'tar': float(tar.detach()), 'grad_norm': gn}
