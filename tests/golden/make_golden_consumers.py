"""Golden fixtures for the callers either side of the hot path, made by RUNNING THE REFERENCE'S PYTHON.

Run in the build container only (needs /root/reference):    python tests/golden/make_golden_consumers.py

Executed unmodified from /root/reference (imported with `mdqe`, `mdqe.models`, `mdqe.util` pre-registered as bare namespace
packages so that mdqe/__init__.py -- detectron2 -- never runs):
  * mdqe/models/matcher.py:11-28   batch_dice_loss        } on out_masks = einsum('bqm,bmthw->bqthw') as in matcher.py:182-195
  * mdqe/models/matcher.py:36-61   batch_sigmoid_ce_loss  }
  * mdqe/util/misc.py:485-507      aligned_bilinear (+ .sigmoid() as in mdqe/mdqe.py:357)
  * mdqe/models/criterion.py:20-43, 51-81, 87-108, 116-145  dice_loss, interinst_dice_loss, sigmoid_ce_loss, interinst_sigmoid_ce_loss
                                   applied to src_masks = einsum(...)[idx] as in loss_masks (:440, :467-473), with autograd for the
                                   gradients (the module's detectron2.projects.point_rend import is satisfied by a stand-in function
                                   that the loss functions never call)
  * mdqe/tracking/OverTracker.py:92-113  OverTracker._get_siou (the module's `from detectron2.structures import Instances` is
                                   satisfied by an empty stand-in class; _get_siou never touches it)
mdqe/mdqe.py (inference_clip, :386-393) and mdqe/models/transformer_dec.py (:167-179) sit inside methods of classes that need
detectron2 / a full model to be constructed; the LINES THEMSELVES are read from the reference files at generation time and executed
(`reference_source`), with the handful of local names they use bound here -- no statement of the reference is re-typed.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def import_reference():
    for name, sub in (("mdqe", "mdqe"), ("mdqe.models", "mdqe/models"), ("mdqe.util", "mdqe/util")):
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = mod
    import mdqe.models.matcher as matcher
    import mdqe.util.misc as misc
    for name in ("detectron2", "detectron2.structures"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["detectron2.structures"].Instances = type("Instances", (), {})
    trk = types.ModuleType("mdqe.tracking")
    trk.__path__ = [os.path.join(REF, "mdqe/tracking")]
    sys.modules["mdqe.tracking"] = trk
    import mdqe.tracking.OverTracker as tracker
    for name in ("detectron2.projects", "detectron2.projects.point_rend", "detectron2.projects.point_rend.point_features"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["detectron2.projects.point_rend.point_features"].get_uncertain_point_coords_on_grid = lambda *a, **k: None
    import mdqe.models.criterion as criterion
    return matcher, misc, tracker, criterion


def reference_source(relpath, first_marker, last_marker):
    """the reference's own source text from the line containing `first_marker` to the line containing `last_marker`, dedented --
    executed, never copied into the repo.  -> (source, first line number, last line number)"""
    import textwrap
    with open(os.path.join(REF, relpath)) as f:
        lines = f.readlines()
    first = next(i for i, l in enumerate(lines) if first_marker in l)
    last = next(i for i, l in enumerate(lines) if last_marker in l and i >= first)
    return textwrap.dedent("".join(lines[first:last + 1])), first + 1, last + 1


def save(name, **arrays):
    conv = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    matcher, misc, tracker, criterion = import_reference()
    g = torch.Generator().manual_seed(11)

    # ---- matcher mask costs, one clip (matcher.py:182, :193-197)
    for name, (Q, K, G, T, H, W) in {"match_cost_K32": (37, 32, 5, 2, 12, 20), "match_cost_K24": (20, 24, 17, 3, 6, 10)}.items():
        coeff = torch.tanh(torch.randn(1, Q, K, generator=g))
        proto = torch.randn(1, K, T, H, W, generator=g)
        tgt = (torch.rand(G, T, H, W, generator=g) > 0.7).float()
        out_masks = torch.einsum('bqm, bmthw -> bqthw', coeff, proto)[0]
        cost_bce = matcher.batch_sigmoid_ce_loss(out_masks, tgt)
        cost_dice = matcher.batch_dice_loss(out_masks, tgt)
        save(name, coeff=coeff[0], proto=proto[0], targets=tgt, cost_bce=cost_bce, cost_dice=cost_dice)

    # ---- NMS soft IoU of inference_clip (mdqe/mdqe.py:394-401), statements copied as they are executed there
    for name, (Q, T, H, W) in {"nms_siou_T4": (23, 4, 12, 20), "nms_siou_T5": (9, 5, 11, 15)}.items():
        mask_pred = torch.randn(Q, T, H, W, generator=g) * 2 - 0.5
        # mdqe/mdqe.py:394-401 run from the reference file itself (the lines between "# Just avoid running out of memory" and the
        # siou matrix of MDQE.inference_clip); checked below against the statements typed out, which is what the first version of this
        # script executed
        src, l0, l1 = reference_source("mdqe/mdqe.py", "mask_nms = mask_pred[:, ::2]", "siou = numerator / (denominator + 1)")
        assert (l0, l1) == (394, 401), f"reference lines moved: {l0}-{l1}"
        env = {"torch": torch, "F": F, "mask_pred": mask_pred}
        exec(compile(src, f"/root/reference/mdqe/mdqe.py:{l0}-{l1}", "exec"), env)
        siou = env["siou"]
        mask_soft = F.interpolate(mask_pred[:, ::2] if T >= 5 else mask_pred, scale_factor=0.5).flatten(1).sigmoid()
        mask_hard = mask_soft.gt(0.5).float()
        numerator = torch.mm(mask_soft, mask_hard.t())
        assert torch.equal(siou, numerator / (mask_soft.sum(-1)[:, None] + mask_hard.sum(-1)[None] - numerator + 1))
        save(name, mask_pred=mask_pred, siou=siou)

    # ---- criterion mask losses of the matched queries (criterion.py:440, :467-473), inter-instance and plain forms
    for name, (Q, K, G, T, H, W, inter) in {"mask_losses_interinst": (30, 32, 6, 2, 12, 20, True), "mask_losses_plain": (12, 24, 3, 3, 6, 10, False)}.items():
        coeff = torch.tanh(torch.randn(1, Q, K, generator=g)).requires_grad_(True)
        proto = torch.randn(1, K, T, H, W, generator=g).requires_grad_(True)
        src_idx = torch.randperm(Q, generator=g)[:G]
        idx = (torch.zeros(G, dtype=torch.long), src_idx)
        tgt = (torch.rand(G, T, H, W, generator=g) > 0.7).float()
        tgt_inter = (torch.rand(G, T, H, W, generator=g) > 0.6).float()
        num_masks = 4.0
        src_masks = torch.einsum('bqm, bmthw -> bqthw', coeff, proto)[idx]                        # criterion.py:440
        if inter:
            loss_mask = criterion.interinst_sigmoid_ce_loss(src_masks, tgt, tgt_inter, num_masks)    # :468
            loss_dice = criterion.interinst_dice_loss(src_masks, tgt, tgt_inter, num_masks)          # :469
        else:
            loss_mask = criterion.sigmoid_ce_loss(src_masks, tgt, num_masks)                         # :474
            loss_dice = criterion.dice_loss(src_masks, tgt, num_masks)                               # :475
        gw = torch.tensor([0.7, 1.9])
        (gw[0] * loss_mask + gw[1] * loss_dice).backward()
        save(name, coeff=coeff[0], proto=proto[0], src_idx=src_idx, targets=tgt, targets_interinst=tgt_inter, interinst=int(inter),
             num_masks=num_masks, grad_weights=gw, loss_mask=loss_mask, loss_dice=loss_dice, grad_coeff=coeff.grad[0], grad_proto=proto.grad[0])

    # ---- tracker mask IoU (OverTracker._get_siou, OverTracker.py:92-113); an empty saved mask and an empty input mask included
    for name, (Ns, Ni, T, H, W) in {"track_siou_a": (7, 5, 3, 10, 14), "track_siou_b": (3, 11, 1, 9, 9)}.items():
        saved = torch.rand(Ns, T, H, W, generator=g)
        inp = torch.rand(Ni, T, H, W, generator=g)
        saved[1] = 0.2
        inp[0] = 0.0
        siou = tracker.OverTracker._get_siou(None, saved, inp)
        save(name, saved_masks=saved, input_masks=inp, siou=siou)

    # ---- aligned_bilinear(pred_masks, factor=match_stride).sigmoid() (mdqe/mdqe.py:357, misc.py:485-507)
    for name, (n, C, H, W, f) in {"aligned_bilinear_f4": (2, 3, 6, 10, 4), "aligned_bilinear_f2": (1, 2, 5, 7, 2)}.items():
        x = torch.randn(n, C, H, W, generator=g) * 3
        up = misc.aligned_bilinear(x, factor=f)
        save(name, x=x, up=up, up_sigmoid=up.sigmoid(), factor=f)

    # ---- query initialisation sampling (transformer_dec.py:170-179), statements as executed there (rearrange == view/permute)
    from einops import rearrange
    B, C = 3, 16
    spatial_shapes = [(6, 10), (3, 5), (2, 3)]
    S = sum(h * w for h, w in spatial_shapes)
    lvl_start_index = [0]
    for h, w in spatial_shapes:
        lvl_start_index.append(lvl_start_index[-1] + h * w)
    encoded_feat = torch.randn(B, S, C, generator=g, dtype=torch.float64).requires_grad_(True)
    n_query_bins = 4
    query_init_coords = (torch.rand(B, n_query_bins * n_query_bins, 2, generator=g, dtype=torch.float64) * 1.1 - 0.05).requires_grad_(True)
    # transformer_dec.py:167-179 run from the reference file itself: coordinates -> grid in [-1, 1], F.grid_sample per pyramid level
    # (bilinear, border padding, align_corners=False), mean over the levels.  `self.n_query_bins` is the only attribute the lines touch.
    src, l0, l1 = reference_source("mdqe/models/transformer_dec.py", "query_init_coords_grid = rearrange(query_init_coords",
                                   "query_init = rearrange(torch.stack(query_init).mean(0)")
    assert (l0, l1) == (167, 179) and "F.grid_sample" in src, f"reference lines moved: {l0}-{l1}"
    env = {"torch": torch, "F": F, "rearrange": rearrange, "self": types.SimpleNamespace(n_query_bins=n_query_bins),
           "query_init_coords": query_init_coords, "spatial_shapes": spatial_shapes, "encoded_feat": encoded_feat,
           "lvl_start_index": lvl_start_index}
    exec(compile(src, f"/root/reference/mdqe/models/transformer_dec.py:{l0}-{l1}", "exec"), env)
    query_init = env["query_init"]
    grad_out = torch.randn(query_init.shape, generator=g, dtype=torch.float64)
    query_init.backward(grad_out)
    save("query_init_f64", feat=encoded_feat, shapes=torch.tensor(spatial_shapes), level_start=torch.tensor(lvl_start_index[:-1]),
         coords=query_init_coords, out=query_init, grad_out=grad_out, grad_feat=encoded_feat.grad, grad_coords=query_init_coords.grad)


if __name__ == "__main__":
    main()
