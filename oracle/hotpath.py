"""oracle/hotpath.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the disparity hot path as the reference models wire it:
`models/SemStereo.py:273-324` (signed, US3D) and `models/SemStereo_WHU.py:273-324` run
against `models/submodule_.py` (unsigned, WHU).

Inputs (what the out-of-scope 2-D part of the model hands to the path):
  f8_l, f8_r   (B,256,H/8,W/8)   features_left[2], features_right[2] after chal_2   (:264,270)
  f4_l, f4_r   (B,128,H/4,W/4)   features_left[1], features_right[1] after chal_1   (:263,269)
  cf_l, cf_r   (B,32,H/4,W/4)    concat_feature(f4_*)                               (:314-315)
  spx_pred     (B,6,H,W)         spx2 output                                        (:271)
  pred_label   (B,6,H,W)         head_l output                                      (:254)
`p` is a dict keyed by the reference's state_dict names.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops

TOPK = 24  # SemStereo.py:301


def attention_branch(p, f8_l, f8_r, maxdisp, signed=True):
    """SemStereo.py:273-279: norm-gwc volume -> patch -> channel gate -> hourglass_att ->
    classif_att_.  Returns the 1/8-res attention logits (B,1,D8,H/8,W/8) and the volume."""
    m8 = maxdisp // 8
    corr = ops.gwc_volume(f8_l, f8_r, m8, 32, signed=signed, norm=True)
    vol = ops.patch_conv(corr, p)
    vol = ops.channel_att(vol, f8_l, p, "corr_feature_att_8")
    vol = ops.hourglass(vol, p, "hourglass_att", (4, 4, 4))
    return ops.classifier(vol, p, "classif_att_"), corr


def attention_stats(p, cost_att, maxdisp, hw4, signed=True):
    """SemStereo.py:279-287 (WHU :279 uses maxdisp//4 bins): trilinear upsample, softmax over
    disparity, mean, variance, sigmoid(beta + gamma*var)."""
    m4 = maxdisp // 4
    nbins = 2 * m4 if signed else m4
    att = ops.trilinear_upsample(cost_att, (nbins, hw4[0], hw4[1]))
    prob = F.softmax(att.squeeze(1), dim=1)
    mu = ops.disparity_regression(prob, m4, signed)
    var = ops.disparity_variance(prob, m4, mu.unsqueeze(1), signed)
    gate = torch.sigmoid(p["beta"] + p["gamma"] * var)
    return att, mu, gate


def sample_strength(f4_l, f4_r, mu, gate):
    """SemStereo.py:288-293: 5 propagated disparity hypotheses, warp-correlate, soft-select."""
    gate5 = ops.propagation(gate)
    d5 = ops.propagation(mu.unsqueeze(1))
    r_w, l_rep = ops.spatial_transformer_grid(f4_l, f4_r, d5)
    corr5 = (l_rep * r_w).mean(dim=1)
    return torch.softmax(corr5 * gate5, dim=1)


def topk_select(att, strength, maxdisp, signed=True, k=TOPK):
    """SemStereo.py:295-310: neighbour-mixed cost column, top-k bins in ascending index order,
    full-softmax probabilities of the kept bins, and the renormalised expectation."""
    mix = (ops.propagation_prob(att) * strength.unsqueeze(2)).sum(dim=1, keepdim=True)
    prob = F.softmax(mix, dim=2)
    ind_k = ops.topk_desc_stable(prob, k, 2).sort(2)[0]
    att_topk = torch.gather(prob, 2, ind_k)
    disp_topk = ind_k.squeeze(1).float()
    if signed:
        disp_topk = disp_topk - maxdisp // 4
    w = F.softmax(torch.gather(mix, 2, ind_k).squeeze(1), dim=1)
    pred_att = (w * disp_topk).sum(dim=1)
    return dict(mix=mix, prob=prob, ind_k=ind_k, att_topk=att_topk, disp_topk=disp_topk, pred_att=pred_att)


def sparse_concat_volume(cf_l, cf_r, disp_topk, att_topk):
    """concat_volume_generator (SemStereo.py:241-244) times att_topk (:318):
    channels [left 0..31 | right-warped 32..63], each scaled by the kept probability."""
    r_w, l_rep = ops.spatial_transformer_grid(cf_l, cf_r, disp_topk)
    return att_topk * torch.cat((l_rep, r_w), dim=1)


def aggregation_branch(p, volume, f4_l):
    """SemStereo.py:319-322: concat_stem -> channel gate -> hourglass2 -> classif."""
    v = ops.conv3d_bn(volume, p, "concat_stem.conv", "concat_stem.bn", 1, 1, True)
    v = ops.channel_att(v, f4_l, p, "concat_feature_att_4")
    v = ops.hourglass(v, p, "hourglass", (6, 4, 4))
    return ops.classifier(v, p, "classif")


@torch.no_grad()
def forward(p, inp, maxdisp, signed=True, att_weights_only=False, keep=False):
    """The whole path.  Returns a dict; `pred_up` / `pred_att_up` are in 1/4-res disparity
    units exactly as `ssr_upsample` returns them (the model multiplies by 4 at :329-346)."""
    f8_l, f8_r, f4_l, f4_r = inp["f8_l"], inp["f8_r"], inp["f4_l"], inp["f4_r"]
    hw4 = f4_l.shape[-2:]
    out = {}
    cost_att, corr = attention_branch(p, f8_l, f8_r, maxdisp, signed)
    att, mu, gate = attention_stats(p, cost_att, maxdisp, hw4, signed)
    strength = sample_strength(f4_l, f4_r, mu, gate)
    sel = topk_select(att, strength, maxdisp, signed)
    out.update(cost_att=cost_att, pred_att0=mu, var_gate=gate, strength=strength,
               ind_k=sel["ind_k"], att_topk=sel["att_topk"], disp_topk=sel["disp_topk"],
               pred_att=sel["pred_att"], prob=sel["prob"])
    if keep:
        out.update(corr_volume=corr, att_weights=att, mix=sel["mix"])
    out["pred_att_up"] = ops.ssr_upsample(sel["pred_att"].unsqueeze(1), inp["spx_pred"], inp["pred_label"], p)
    if att_weights_only:
        return out
    cf_l = inp["cf_l"] if inp.get("cf_l") is not None else ops.concat_feature(f4_l, p)        # SemStereo.py:314-315
    cf_r = inp["cf_r"] if inp.get("cf_r") is not None else ops.concat_feature(f4_r, p)
    volume = sparse_concat_volume(cf_l, cf_r, sel["disp_topk"], sel["att_topk"])
    cost = aggregation_branch(p, volume, f4_l)
    pred = ops.regression_topk(cost.squeeze(1), sel["disp_topk"], 2)
    out.update(cost=cost, pred=pred)
    if keep:
        out["volume"] = volume
    out["pred_up"] = ops.ssr_upsample(pred, inp["spx_pred"], inp["pred_label"], p)
    return out
