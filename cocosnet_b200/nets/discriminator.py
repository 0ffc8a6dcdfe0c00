"""Multiscale PatchGAN discriminator (reference
models/networks/discriminator.py:16-177)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import Attention, BaseNetwork, equal_lr, nonspade_norm, norm_act


class MultiscaleDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument("--netD_subarch", type=str, default="n_layer")
        parser.add_argument("--num_D", type=int, default=2)
        NLayerDiscriminator.modify_commandline_options(parser, is_train)
        return parser

    def __init__(self, opt, stage1=False):
        super().__init__()
        self.opt = opt
        self.stage1 = stage1
        if opt.netD_subarch != "n_layer":
            raise ValueError("unrecognized discriminator subarchitecture %s" % opt.netD_subarch)
        for i in range(opt.num_D):
            self.add_module("discriminator_%d" % i, NLayerDiscriminator(opt, stage1=stage1))

    @staticmethod
    def downsample(x):
        return F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)

    def forward(self, input):
        result, segs, cam_logits = [], [], []
        keep_feats = not self.opt.no_ganFeat_loss
        for _, D in self.named_children():
            out, cam = D(input)
            cam_logits.append(cam)
            result.append(out if keep_feats else [out])
            input = self.downsample(input)
        return result, segs, cam_logits


class NLayerDiscriminator(BaseNetwork):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument("--n_layers_D", type=int, default=4)
        return parser

    def __init__(self, opt, stage1=False):
        super().__init__()
        self.opt = opt
        self.stage1 = stage1
        kw, padw, nf = 4, 1, opt.ndf
        input_nc = opt.label_nc + opt.output_nc + (1 if opt.contain_dontcare_label else 0)
        norm = nonspade_norm(opt, opt.norm_D)
        seq = [[nn.Conv2d(input_nc, nf, kernel_size=kw, stride=2, padding=padw), nn.LeakyReLU(0.2, False)]]
        self.use_attn = ((not stage1) and opt.use_attention) or (stage1 and getattr(opt, "use_attention_st1", False))
        for n in range(1, opt.n_layers_D):
            nf_prev, nf = nf, min(nf * 2, 512)
            last = n == opt.n_layers_D - 1
            if self.use_attn and last:
                self.attn = Attention(nf_prev, "spectral" in opt.norm_D)
            if last and not stage1:
                # built by the reference but never used in forward (discriminator.py:101-110);
                # kept so state_dict keys / checkpoints stay compatible
                dec, nc = [], nf_prev
                for _ in range(opt.n_layers_D - 1):
                    dec += [nn.Upsample(scale_factor=2),
                            norm(nn.Conv2d(nc, nc // 2, kernel_size=3, stride=1, padding=1)), nn.LeakyReLU(0.2, False)]
                    nc //= 2
                dec += [nn.Conv2d(nc, opt.semantic_nc, kernel_size=3, stride=1, padding=1)]
                self.dec = nn.Sequential(*dec)
            seq += [[norm(nn.Conv2d(nf_prev, nf, kernel_size=kw, stride=1 if last else 2, padding=padw)),
                     nn.LeakyReLU(0.2, False)]]
        seq += [[nn.Conv2d(nf, 1, kernel_size=kw, stride=1, padding=padw)]]
        if opt.D_cam > 0:
            mult = min(2 ** (opt.n_layers_D - 1), 8)
            wrap = equal_lr if opt.eqlr_sn else nn.utils.spectral_norm
            self.gap_fc = wrap(nn.Linear(opt.ndf * mult, 1, bias=False))
            self.gmp_fc = wrap(nn.Linear(opt.ndf * mult, 1, bias=False))
            self.conv1x1 = nn.Conv2d(opt.ndf * mult * 2, opt.ndf * mult, kernel_size=1, stride=1, bias=True)
            self.leaky_relu = nn.LeakyReLU(0.2, True)
        for n, layers in enumerate(seq):
            self.add_module("model" + str(n), nn.Sequential(*layers))

    def forward(self, input):
        results = [input]
        cam_logit = None
        for name, sub in self.named_children():
            if "model" not in name:
                continue
            x = results[-1]
            if name == "model3" and self.use_attn:
                x = self.attn(x)
            if len(sub) == 2 and isinstance(sub[1], nn.LeakyReLU):  # [conv (+ norm), LeakyReLU]
                y = norm_act(sub[0], x, sub[1].negative_slope)
            else:
                y = sub(x)
            if self.opt.D_cam > 0 and name == "model3":
                gap = F.adaptive_avg_pool2d(y, 1)
                gap_logit = self.gap_fc(gap.view(y.shape[0], -1))
                gap = y * list(self.gap_fc.parameters())[0].unsqueeze(2).unsqueeze(3)
                gmp = F.adaptive_max_pool2d(y, 1)
                gmp_logit = self.gmp_fc(gmp.view(y.shape[0], -1))
                gmp = y * list(self.gmp_fc.parameters())[0].unsqueeze(2).unsqueeze(3)
                cam_logit = torch.cat([gap_logit, gmp_logit], 1)
                y = self.leaky_relu(self.conv1x1(torch.cat([gap, gmp], 1)))
            results.append(y)
        return (results[1:] if not self.opt.no_ganFeat_loss else results[-1]), cam_logit
