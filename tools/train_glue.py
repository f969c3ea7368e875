"""SemStereoTrainGlue -- a stand-in for the reference model file on machines where the reference tree is absent (the GPU box).

It restates `SemStereo.__init__` / `forward` (models/SemStereo.py:184-346; training-mode returns :329-337) for
attention_weights_only=True the way the reference writes them: the STAR-IMPORTED names come from the operator surface
(`semstereo_b200.submodule`, `submodule_other` -- kernel-backed and differentiable, semstereo_b200/train_ops.py, torch_ops.py),
what the reference does inline is plain torch (nn.ConvTranspose3d, the `patch` / classifier nn.Conv3d, F.interpolate, softmax,
sort, gather), and so are the 2-D decoder modules (Conv2x, segmenthead, 2-D BasicConv: torch modules of the surface).  The backbone
stays the caller's (`Feature` = timm): the glue starts from its five feature maps.  Used by tools/train_step.py (BASELINE config
#5) and tests/test_gpu_train.py; with the reference tree available one trains the reference model object itself instead
(INTEGRATION.md level 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from semstereo_b200.params import CHANS, CHANS2  # noqa: E402


def _surface():
    from semstereo_b200 import submodule as s, submodule_other as o
    return s, o


class SemStereoTrainGlue(nn.Module):
    """SemStereo(maxdisp, attention_weights_only=True, seg_if=True, stereo_if=True, num_classes) after the backbone, with the
    reference's parameter names (state_dict-compatible with the reference model minus `feature.*` and the aggregation branch's
    modules, which attention_weights_only never runs)."""

    def __init__(self, maxdisp: int = 64, num_classes: int = 6):
        super().__init__()
        s, o = _surface()
        self.maxdisp, self.num_classes = maxdisp, num_classes
        bc = (64, 128, 256, 384, 512)

        class FeatUp(nn.Module):
            def __init__(self):
                super().__init__()
                self.deconv32_16 = s.Conv2x(bc[4], bc[3], deconv=True, concat=True)
                self.deconv16_8 = s.Conv2x(bc[3] * 2, bc[2], deconv=True, concat=True)
                self.deconv8_4 = s.Conv2x(bc[2] * 2, bc[1], deconv=True, concat=True)
                self.deconv4_2 = s.Conv2x(bc[1] * 2, bc[0], deconv=True, concat=True)

            def forward(self, f):
                x2, x4, x8, x16, x32 = f
                x16 = self.deconv32_16(x32, x16)
                x8 = self.deconv16_8(x16, x8)
                x4 = self.deconv8_4(x8, x4)
                x2 = self.deconv4_2(x4, x2)
                return [x2, x4, x8, x16, x32]

        class channelAtt(nn.Module):
            def __init__(self, cv, im):
                super().__init__()
                self.im_att = nn.Sequential(s.BasicConv(im, im // 2, kernel_size=1, stride=1, padding=0), nn.Conv2d(im // 2, cv, 1))

            def forward(self, cv, im):
                return torch.sigmoid(self.im_att(im).unsqueeze(2)) * cv

        class hourglass(nn.Module):
            def __init__(self, c, block):
                super().__init__()
                self.conv1 = nn.Sequential(o.convbn_3d(c, c * 2, 3, 2, 1), nn.ReLU(inplace=True))
                self.conv2 = nn.Sequential(o.convbn_3d(c * 2, c * 2, 3, 1, 1), nn.ReLU(inplace=True))
                self.conv3 = nn.Sequential(o.convbn_3d(c * 2, c * 4, 3, 2, 1), nn.ReLU(inplace=True))
                self.conv4 = nn.Sequential(o.convbn_3d(c * 4, c * 4, 3, 1, 1), nn.ReLU(inplace=True))
                self.attention_block = o.attention_block(channels_3d=c * 4, num_heads=16, block=block)
                self.conv5 = nn.Sequential(nn.ConvTranspose3d(c * 4, c * 2, 3, padding=1, output_padding=1, stride=2, bias=False), nn.BatchNorm3d(c * 2))
                self.conv6 = nn.Sequential(nn.ConvTranspose3d(c * 2, c, 3, padding=1, output_padding=1, stride=2, bias=False), nn.BatchNorm3d(c))
                self.redir1 = o.convbn_3d(c, c, kernel_size=1, stride=1, pad=0)
                self.redir2 = o.convbn_3d(c * 2, c * 2, kernel_size=1, stride=1, pad=0)

            def forward(self, x):
                c1 = self.conv1(x)
                c2 = self.conv2(c1)
                c4 = self.attention_block(self.conv4(self.conv3(c2)))
                c5 = F.relu(self.conv5(c4) + self.redir2(c2), inplace=True)
                return F.relu(self.conv6(c5) + self.redir1(x), inplace=True)

        self.feature_up = FeatUp()
        self.head_l = s.segmenthead(CHANS[0], CHANS[0] // 4, num_classes, scale_factor=2)
        self.head_r = s.segmenthead(CHANS[0], CHANS[0] // 4, num_classes, scale_factor=2)
        self.gamma, self.beta = nn.Parameter(torch.zeros(1)), nn.Parameter(2 * torch.ones(1))
        self.spx2 = nn.Sequential(nn.ConvTranspose2d(CHANS2[0] * 2, 6, kernel_size=4, stride=2, padding=1))
        self.spx4_2 = s.Conv2x(CHANS2[1] * 2, CHANS2[0], True)
        self.spx8_4 = s.Conv2x(CHANS2[2] * 2, CHANS2[1], True)
        self.spx16_8 = s.Conv2x(CHANS2[3] * 2, CHANS2[2], True)
        self.spx32_16 = s.Conv2x(CHANS2[4], CHANS2[3], True)
        for i in range(5):
            setattr(self, f"chal_{i}", nn.Sequential(nn.Conv2d(CHANS[i], CHANS2[i], kernel_size=1, stride=1), nn.BatchNorm2d(CHANS2[i])))
        self.patch = nn.Conv3d(32, 32, kernel_size=(1, 3, 3), stride=1, groups=32, padding=(0, 1, 1), bias=False)
        self.corr_feature_att_8 = channelAtt(32, CHANS2[2])
        self.hourglass_att = hourglass(32, (4, 4, 4))
        self.classif_att_ = nn.Sequential(o.convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv3d(32, 1, 3, padding=1, bias=False))
        self.propagation, self.propagation_prob = s.Propagation(), s.Propagation_prob()
        self.ssr_upsample = s.SSR_upsample(num_classes)
        self._s = [s]

    def forward(self, feat_l, feat_r):
        """feat_l / feat_r: the backbone's five maps per image.  Training mode returns what SemStereo.forward does (:329-332):
        ([pred_att_up*4, pred_att*4], pred_label, pred_label_r); eval mode ([pred_att_up*4], pred_label)."""
        s = self._s[0]
        md = self.maxdisp
        fl, fr = self.feature_up(list(feat_l)), self.feature_up(list(feat_r))
        pred_label, pred_label_r = self.head_l(fl[0]), self.head_r(fr[0])
        fl = [getattr(self, f"chal_{i}")(fl[i]) for i in range(5)]
        fr1, fr2 = self.chal_1(fr[1]), self.chal_2(fr[2])
        x = self.spx32_16(fl[4], fl[3])
        x = self.spx16_8(x, fl[2])
        x = self.spx8_4(x, fl[1])
        x = self.spx4_2(x, fl[0])
        spx_pred = self.spx2(x)
        corr = self.patch(s.build_gwc_volume_norm(fl[2], fr2, md // 8, 32))
        cost_att = self.classif_att_(self.hourglass_att(self.corr_feature_att_8(corr, fl[2])))
        att = F.interpolate(cost_att, [md // 4 * 2, fl[1].shape[2], fl[1].shape[3]], mode="trilinear")
        prob = F.softmax(att.squeeze(1), dim=1)
        mu = s.disparity_regression(prob, md // 4)
        var = torch.sigmoid(self.beta + self.gamma * s.disparity_variance(prob, md // 4, mu.unsqueeze(1)))
        var5, d5 = self.propagation(var), self.propagation(mu.unsqueeze(1))
        r_w, l_rep = s.SpatialTransformer_grid(fl[1], fr1, d5)
        strength = torch.softmax((l_rep * r_w).mean(dim=1) * var5, dim=1)
        mix = torch.sum(self.propagation_prob(att) * strength.unsqueeze(2), dim=1, keepdim=True)
        p = F.softmax(mix, dim=2)
        ind_k = p.sort(2, True)[1][:, :, :24].sort(2, False)[0]
        samples = ind_k.squeeze(1).float() - md // 4
        w = F.softmax(torch.gather(mix, 2, ind_k).squeeze(1), dim=1)
        pred_att = torch.sum(w * samples, dim=1)
        pred_att_up = self.ssr_upsample(pred_att.unsqueeze(1), spx_pred, pred_label)
        if self.training:
            return [pred_att_up * 4, pred_att * 4], pred_label, pred_label_r
        return [pred_att_up * 4], pred_label


def synthetic_batch(seed, B, H, W, device, maxdisp=64, num_classes=6):
    """Config #5's synthetic targets: disp_gt ~ U[-maxdisp, maxdisp), label ~ randint(0, num_classes); the backbone
    pyramids come from params.make_backbone_features."""
    from semstereo_b200.params import make_backbone_features
    g = torch.Generator().manual_seed(seed)
    fl, fr = make_backbone_features(seed, B, H, W)
    disp = (torch.rand(B, H, W, generator=g) * 2 - 1) * maxdisp
    disp4 = F.interpolate(disp.unsqueeze(1), scale_factor=0.25, mode="nearest").squeeze(1)
    label = torch.randint(0, num_classes, (B, H, W), generator=g).float()
    mv = lambda t: t.to(device)      # noqa: E731
    return [mv(t) for t in fl], [mv(t) for t in fr], mv(disp), mv(disp4), mv(label)
