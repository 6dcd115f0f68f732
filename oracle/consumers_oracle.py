"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the callers either side of the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product package never does.
Pinned against fixtures produced by the reference's own Python (tests/golden/make_golden_consumers.py,
tests/test_consumers_oracle.py).  Paths relative to /root/reference.
"""
import numpy as np


def _softplus(x):
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x):
    e = np.exp(-np.abs(x))
    return np.where(x >= 0, 1.0, e) / (1.0 + e)


def match_cost(coeff, proto, targets):
    """mdqe/models/matcher.py:182 (out_masks = einsum), :36-61 (batch_sigmoid_ce_loss), :11-28 (batch_dice_loss), one clip.
    coeff [Q,K], proto [K,...], targets [G,...] -> cost_bce [Q,G], cost_dice [Q,G]."""
    K = coeff.shape[1]
    p = np.asarray(proto, dtype=np.float64).reshape(K, -1)
    t = np.asarray(targets, dtype=np.float64).reshape(targets.shape[0], -1)
    x = np.asarray(coeff, dtype=np.float64) @ p                     # matcher.py:182
    pos, neg = _softplus(-x), _softplus(x)                          # BCE-with-logits against ones / zeros (:51-56)
    cost_bce = (pos @ t.T + neg @ (1 - t).T) / x.shape[1]          # :58-61
    s = _sigmoid(x)                                                 # :22
    cost_dice = 1 - (2 * (s @ t.T) + 1) / (s.sum(-1)[:, None] + t.sum(-1)[None, :] + 1)   # :25-27
    return cost_bce, cost_dice


def mask_losses(coeff, proto, targets, targets_interinst, num_masks, grad_weights=None):
    """mdqe/models/criterion.py:440-473 for the matched rows: coeff [G,K], proto [K,...], targets / targets_interinst [G,...]
    (targets_interinst None: sigmoid_ce_loss :87-108 + dice_loss :20-43, else the inter-instance forms :116-145, :51-81).
    Returns (loss_mask, loss_dice) and, with grad_weights = (d/d loss_mask, d/d loss_dice), also (grad_coeff, grad_proto)."""
    G, K = coeff.shape
    c = np.asarray(coeff, dtype=np.float64)
    p = np.asarray(proto, dtype=np.float64).reshape(K, -1)
    t = np.asarray(targets, dtype=np.float64).reshape(G, -1)
    x = c @ p                                                                                     # :440
    s = _sigmoid(x)
    l = _softplus(x) - x * t                                                                      # binary_cross_entropy_with_logits
    inv = 1.0 / max(num_masks, 1)
    if targets_interinst is None:
        w = np.ones_like(x)
        wsum = np.full(G, float(x.shape[1]))                                                      # loss.mean(1) (:108)
        tib = np.zeros_like(x)
    else:
        ti = np.asarray(targets_interinst, dtype=np.float64).reshape(G, -1)
        w = ti + 1                                                                                # :140
        wsum = np.maximum(w.sum(1), 1)                                                            # :142
        tib = ((ti > 0.5) & ((1 - t) > 0.5)).astype(np.float64)                                   # :69
    loss_mask = ((l * w).sum(1) / wsum).sum() * inv                                               # :142-144
    num = 2 * (s * t).sum(1) + ((1 - s) * tib).sum(1)                                             # :77 (:39 without tib)
    den = s.sum(1) + t.sum(1) + tib.sum(1)                                                        # :78
    loss_dice = (1 - (num + 1) / (den + 1)).sum() * inv                                           # :79-81
    if grad_weights is None:
        return loss_mask, loss_dice
    gm, gd = float(grad_weights[0]) * inv, float(grad_weights[1]) * inv
    ds = s * (1 - s)
    gx = gm * (s - t) * w / wsum[:, None] + gd * ds * ((num + 1)[:, None] / (den + 1)[:, None] ** 2 - (2 * t - tib) / (den + 1)[:, None])
    return loss_mask, loss_dice, gx @ p.T, (c.T @ gx).reshape(np.asarray(proto).shape)


def nms_siou(mask_pred):
    """mdqe/mdqe.py:394-401.  mask_pred [Q,T,H,W] -> siou [Q,Q]."""
    m = np.asarray(mask_pred)
    nms = m[:, ::2] if m.shape[1] >= 5 else m                       # :386
    H2, W2 = nms.shape[2] // 2, nms.shape[3] // 2
    nms = nms[:, :, 0:2 * H2:2, 0:2 * W2:2]                         # F.interpolate(scale_factor=0.5), nearest: source pixel 2*i (:387)
    soft = _sigmoid(nms.reshape(m.shape[0], -1).astype(np.float32)).astype(np.float32)
    hard = (soft > 0.5).astype(np.float64)                          # :388
    soft = soft.astype(np.float64)
    num = soft @ hard.T                                             # :391
    den = soft.sum(-1)[:, None] + hard.sum(-1)[None, :] - num       # :392
    return num / (den + 1)                                          # :393


def track_siou(saved_masks, input_masks):
    """mdqe/tracking/OverTracker.py:92-113 (OverTracker._get_siou).  [Ns,T,H,W], [Ni,T,H,W] -> siou [Ns,Ni]."""
    i = (np.asarray(input_masks).reshape(input_masks.shape[0], -1) > 0.5).astype(np.float64)      # :98
    s = (np.asarray(saved_masks).reshape(saved_masks.shape[0], -1) > 0.5).astype(np.float64)      # :99
    valid = (s.any(-1)[:, None] & i.any(-1)[None, :]).astype(np.float64)                          # :103
    num = s @ i.T                                                                                 # :106
    den = s.sum(-1)[:, None] + i.sum(-1)[None, :] - num                                           # :107
    return (num * valid) / (den * valid + 1e-6)                                                   # :109-111


def aligned_bilinear(x, factor):
    """mdqe/util/misc.py:485-507.  x [..., H, W] -> [..., factor*H, factor*W]."""
    x = np.asarray(x)
    if factor == 1:
        return x
    h, w = x.shape[-2:]
    x = np.concatenate([x, x[..., -1:, :]], -2)                     # F.pad(0, 1, 0, 1, replicate) (:494)
    x = np.concatenate([x, x[..., :, -1:]], -1)
    oh, ow = factor * h + 1, factor * w + 1                         # align_corners=True resize (:495-501): source = dst * h / (factor * h)
    sy = np.arange(oh, dtype=np.float64) * (h / (oh - 1))
    sx = np.arange(ow, dtype=np.float64) * (w / (ow - 1))
    y0 = np.minimum(np.floor(sy).astype(int), h)
    x0 = np.minimum(np.floor(sx).astype(int), w)
    y1, x1 = np.minimum(y0 + 1, h), np.minimum(x0 + 1, w)
    ly, lx = (sy - y0)[:, None], (sx - x0)[None, :]
    top = x[..., y0, :][..., :, x0] * (1 - lx) + x[..., y0, :][..., :, x1] * lx
    bot = x[..., y1, :][..., :, x0] * (1 - lx) + x[..., y1, :][..., :, x1] * lx
    up = top * (1 - ly) + bot * ly
    p = factor // 2                                                 # F.pad(p, 0, p, 0, replicate) (:502-505), crop (:507)
    up = np.concatenate([np.repeat(up[..., :1, :], p, -2), up], -2)
    up = np.concatenate([np.repeat(up[..., :, :1], p, -1), up], -1)
    return up[..., :oh - 1, :ow - 1]


def _border(c, size):
    """grid_sample(padding_mode='border', align_corners=False) for a normalised coordinate c (grid = 2c - 1): pixel coordinate
    clipped to [0, size-1], its derivative w.r.t. c (0 where clipped), the two corner indices and the fraction."""
    x = c * size - 0.5
    g = np.where((x > 0) & (x < size - 1), float(size), 0.0)
    x = np.clip(x, 0, size - 1)
    x0 = np.floor(x).astype(int)
    return x0, x0 + 1, x - x0, g


def query_init_sample(feat, shapes, level_start, coords, grad_out=None):
    """mdqe/models/transformer_dec.py:170-179: per level F.grid_sample(bilinear, border, align_corners=False), mean over levels.
    feat [B,S,C], coords [B,Q,2] (x, y) normalised -> out [B,Q,C]; with grad_out also (grad_feat, grad_coords)."""
    feat = np.asarray(feat, dtype=np.float64)
    coords = np.asarray(coords, dtype=np.float64)
    B, S, C = feat.shape
    Q = coords.shape[1]
    L = len(shapes)
    out = np.zeros((B, Q, C))
    gf = np.zeros_like(feat)
    gc = np.zeros_like(coords)
    for l in range(L):
        H, W = int(shapes[l][0]), int(shapes[l][1])
        st = int(level_start[l])
        x0, x1, lx, gx = _border(coords[..., 0], W)
        y0, y1, ly, gy = _border(coords[..., 1], H)
        for b in range(B):
            f = feat[b, st:st + H * W].reshape(H, W, C)
            for q in range(Q):
                acc = np.zeros(C)
                dx = np.zeros(C)
                dy = np.zeros(C)
                for (yy, wy, sy) in ((y0[b, q], 1 - ly[b, q], -1.0), (y1[b, q], ly[b, q], 1.0)):
                    for (xx, wx, sx) in ((x0[b, q], 1 - lx[b, q], -1.0), (x1[b, q], lx[b, q], 1.0)):
                        if 0 <= yy < H and 0 <= xx < W:
                            acc += wy * wx * f[yy, xx]
                            dx += sx * wy * f[yy, xx]
                            dy += sy * wx * f[yy, xx]
                            if grad_out is not None:
                                gf[b, st + yy * W + xx] += wy * wx * grad_out[b, q] / L
                out[b, q] += acc / L
                if grad_out is not None:
                    gc[b, q, 0] += gx[b, q] * np.dot(dx, grad_out[b, q]) / L
                    gc[b, q, 1] += gy[b, q] * np.dot(dy, grad_out[b, q]) / L
    if grad_out is None:
        return out
    return out, gf, gc
