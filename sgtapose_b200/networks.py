"""Network modules with the reference's parameter names and forward signature.

`create_model('dlapawdl3new_34', heads, head_conv, opt)` (reference
sgtapose/lib/model/model.py:24-29) returns `DLA_PlanAWindow_l3new`
(sgtapose/lib/model/networks/dla.py:1458-1554 on top of BaseModelPlanA,
networks/base_model.py:102-200):

    model(x, pre_img, pre_hm, repro_hm, pre_hm_cls, repro_hm_cls) -> [ {hm, reg, tracking} ]

State-dict keys are identical to the reference's (base.*, dla_up.ida_*.{proj,up,node}_*,
ida_up.*, transformer.{0,1,2}.layers.{0,1,2}.* -- three aliases of ONE shared layer,
dla.py:788-789 --, cat_layer.*, hm/reg/tracking), so reference checkpoints load unchanged.

This module tree is the differentiable (training / eager) form: plain convolutions go
through cuDNN via torch, while DCN, the attention core and the token path run in
libsgta_b200.so.  `engine.InferenceEngine` compiles the same parameters into the fused
NHWC inference path.
"""
import math

import numpy as np
import torch
from torch import nn

from . import fusion
from .dcn_v2 import DCN

BN_MOMENTUM = 0.1
DLA34_LEVELS = (1, 1, 1, 2, 2, 1)
DLA34_CHANNELS = (16, 32, 64, 128, 256, 512)


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


def _stem(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 7, 1, 3, bias=False), _bn(cout), nn.ReLU(inplace=True))


class BasicBlock(nn.Module):
    """conv-bn-relu-conv-bn (+residual) relu  (dla.py:41-69)."""

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = _bn(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = _bn(cout)

    def forward(self, x, residual=None):
        residual = x if residual is None else residual
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + residual)


class Root(nn.Module):
    """1x1 conv over the channel concat of the children (dla.py:157-175)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, 1, 0, bias=False)
        self.bn = _bn(cout)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, *xs):
        return self.relu(self.bn(self.conv(torch.cat(xs, 1))))


class Tree(nn.Module):
    """Hierarchical aggregation node (dla.py:178-231), BasicBlock leaves only."""

    def __init__(self, levels, cin, cout, stride=1, level_root=False, root_dim=0):
        super().__init__()
        root_dim = root_dim or 2 * cout
        if level_root:
            root_dim += cin
        if levels == 1:
            self.tree1 = BasicBlock(cin, cout, stride)
            self.tree2 = BasicBlock(cout, cout, 1)
            self.root = Root(root_dim, cout)
        else:
            self.tree1 = Tree(levels - 1, cin, cout, stride)
            self.tree2 = Tree(levels - 1, cout, cout, root_dim=root_dim + cout)
        self.levels, self.level_root = levels, level_root
        self.downsample = nn.MaxPool2d(stride, stride=stride) if stride > 1 else None
        self.project = None
        if cin != cout:
            self.project = nn.Sequential(nn.Conv2d(cin, cout, 1, 1, bias=False), _bn(cout))

    def forward(self, x, residual=None, children=None):
        children = [] if children is None else children
        bottom = self.downsample(x) if self.downsample is not None else x
        if self.level_root:
            children.append(bottom)
        if self.levels == 1:
            residual = self.project(bottom) if self.project is not None else bottom
            x1 = self.tree1(x, residual)
            x2 = self.tree2(x1)
            return self.root(x2, x1, *children)
        # deeper trees: the reference also evaluates self.project(bottom) here and discards it
        x1 = self.tree1(x)
        children.append(x1)
        return self.tree2(x1, children=children)


class DLA(nn.Module):
    """DLA-34 base with the pre-image / pre-heatmap stems (dla.py:234-337)."""

    def __init__(self, levels=DLA34_LEVELS, channels=DLA34_CHANNELS, opt=None):
        super().__init__()
        self.channels = list(channels)
        c = self.channels
        self.base_layer = _stem(3, c[0])            # never executed on this path, kept for the keys
        self.level0 = self._conv_level(c[0], c[0], levels[0])
        self.level1 = self._conv_level(c[0], c[1], levels[1], stride=2)
        self.level2 = Tree(levels[2], c[1], c[2], 2, level_root=False)
        self.level3 = Tree(levels[3], c[2], c[3], 2, level_root=True)
        self.level4 = Tree(levels[4], c[3], c[4], 2, level_root=True)
        self.level5 = Tree(levels[5], c[4], c[5], 2, level_root=True)
        if opt is None or getattr(opt, "pre_img", True):
            self.pre_img_layer = _stem(3, c[0])
        if opt is None or getattr(opt, "pre_hm", True):
            self.pre_hm_layer = _stem(1, c[0])
        if opt is not None and getattr(opt, "ct_modify", False):
            self.repro_hm_layer = _stem(1, c[0])

    @staticmethod
    def _conv_level(cin, cout, convs, stride=1):
        mods = []
        for i in range(convs):
            mods += [nn.Conv2d(cin, cout, 3, stride if i == 0 else 1, 1, bias=False), _bn(cout),
                     nn.ReLU(inplace=True)]
            cin = cout
        return nn.Sequential(*mods)

    def forward(self, x=None, pre_img=None, pre_hm=None, repro_hm=None):
        if x is not None:
            x = self.base_layer(x)
            if pre_img is not None:
                x = x + self.pre_img_layer(pre_img)
            if pre_hm is not None:
                x = x + self.pre_hm_layer(pre_hm)
            if repro_hm is not None:
                x = x + self.repro_hm_layer(repro_hm)
        elif pre_img is not None:
            x = self.pre_img_layer(pre_img)
            if pre_hm is not None:
                x = x + self.pre_hm_layer(pre_hm)
        elif pre_hm is not None:
            x = self.pre_hm_layer(pre_hm)
        ys = []
        for i in range(6):
            x = getattr(self, "level%d" % i)(x)
            ys.append(x)
        return ys


def fill_up_weights(up):
    """Bilinear initialisation of the depth-wise up-sampler (dla.py:486-495)."""
    w = up.weight.data
    k = w.size(2)
    f = math.ceil(k / 2)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    ramp = torch.tensor([1 - abs(i / f - c) for i in range(k)], dtype=w.dtype)
    w[:, 0] = (ramp[:, None] * ramp[None, :])[None]


class DeformConv(nn.Module):
    """DCN -> BN -> ReLU (dla.py:538-550)."""

    def __init__(self, chi, cho):
        super().__init__()
        self.actf = nn.Sequential(_bn(cho), nn.ReLU(inplace=True))
        self.conv = DCN(chi, cho, kernel_size=(3, 3), stride=1, padding=1, dilation=1,
                        deformable_groups=1)

    def forward(self, x):
        return self.actf(self.conv(x))


class IDAUp(nn.Module):
    """Iterative deep aggregation up-sampling (dla.py:552-577)."""

    def __init__(self, o, channels, up_f):
        super().__init__()
        for i in range(1, len(channels)):
            f = int(up_f[i])
            setattr(self, "proj_%d" % i, DeformConv(channels[i], o))
            up = nn.ConvTranspose2d(o, o, f * 2, stride=f, padding=f // 2, output_padding=0,
                                    groups=o, bias=False)
            fill_up_weights(up)
            setattr(self, "up_%d" % i, up)
            setattr(self, "node_%d" % i, DeformConv(o, o))

    def forward(self, layers, startp, endp):
        for i in range(startp + 1, endp):
            k = i - startp
            layers[i] = getattr(self, "up_%d" % k)(getattr(self, "proj_%d" % k)(layers[i]))
            layers[i] = getattr(self, "node_%d" % k)(layers[i] + layers[i - 1])


class DLAUp(nn.Module):
    """dla.py:581-606."""

    def __init__(self, startp, channels, scales, in_channels=None):
        super().__init__()
        self.startp = startp
        in_channels = list(channels) if in_channels is None else list(in_channels)
        channels = list(channels)
        scales = np.array(scales, dtype=int)
        for i in range(len(channels) - 1):
            j = -i - 2
            setattr(self, "ida_%d" % i, IDAUp(channels[j], in_channels[j:], scales[j:] // scales[j]))
            scales[j + 1:] = scales[j]
            in_channels[j + 1:] = [channels[j] for _ in channels[j + 1:]]

    def forward(self, layers):
        out = [layers[-1]]
        for i in range(len(layers) - self.startp - 1):
            getattr(self, "ida_%d" % i)(layers, len(layers) - i - 2, len(layers))
            out.insert(0, layers[-1])
        return out


class TransformerEncoderLayer(nn.Module):
    """Cross-attention + LN + FFN + LN (dla.py:702-743); dropout modules kept for parity of
    train-mode behaviour."""

    def __init__(self, d_inp, d_model, n_k, d_ffn=1024, dropout=0.1, n_heads=8, pos_embed=True):
        super().__init__()
        self.d_model, self.d_inp, self.d_ffn, self.n_heads = d_model, d_inp, d_ffn, n_heads
        self.d_out = d_model * n_heads
        self.cross_attn = fusion.MHCA_ein(n_heads, d_inp, self.d_out, n_k, pos_embed=pos_embed)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_inp)
        self.linear1 = nn.Linear(d_inp, d_ffn)
        self.activation = nn.ReLU()
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_inp)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_inp)

    def forward(self, query, key, value):
        query = self.norm1(self.cross_attn(query, key, value) + self.dropout1(query))
        ffn = self.linear2(self.dropout3(self.activation(self.linear1(query))))
        return self.norm3(query + self.dropout4(ffn))


class TransformerEncoder(nn.Module):
    """N applications of ONE shared layer (dla.py:788-803: `_get_clones` aliases the module)."""

    def __init__(self, layer, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([layer for _ in range(num_layers)])
        self.num_layers = num_layers

    def forward(self, query, key, value):
        out = query
        for layer in self.layers:
            out = layer(out, key, value)
        return out


def _make_head(last_channel, classes, head_conv, head_kernel, is_hm, prior_bias):
    """base_model.py:121-162."""
    if len(head_conv) > 0:
        convs = [nn.Conv2d(last_channel, head_conv[0], head_kernel, padding=head_kernel // 2, bias=True)]
        for k in range(1, len(head_conv)):
            convs.append(nn.Conv2d(head_conv[k - 1], head_conv[k], 1, bias=True))
        out = nn.Conv2d(head_conv[-1], classes, 1, 1, 0, bias=True)
        mods = []
        for c in convs:
            mods += [c, nn.ReLU(inplace=True)]
        fc = nn.Sequential(*mods, out)
        last = fc[-1]
    else:
        fc = nn.Conv2d(last_channel, classes, 1, 1, 0, bias=True)
        last = fc
    if is_hm:
        last.bias.data.fill_(prior_bias)
    else:
        for m in fc.modules():
            if isinstance(m, nn.Conv2d) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
    return fc


class DLA_PlanAWindow_l3new(nn.Module):
    SCALE_LIST = (4, 2, 1, 1 / 2, 1 / 4, 1 / 8)

    def __init__(self, num_layers, heads, head_convs, opt):
        super().__init__()
        if num_layers != 34:
            raise ValueError("only the shipped DLA-34 configuration is built (arch dlapawdl3new_34)")
        if getattr(opt, "dla_node", "dcn") != "dcn":
            raise ValueError("only --dla_node dcn is built")
        self.opt = opt
        self.heads = heads
        self.num_stacks = 1
        head_kernel = getattr(opt, "head_kernel", 3)
        for head in heads:
            if "wh" in head:                     # base_model.py:114-115
                continue
            setattr(self, head, _make_head(64, heads[head], head_convs[head], head_kernel,
                                           "hm" in head, getattr(opt, "prior_bias", -4.6)))
        self.first_level, self.last_level = 2, 5
        self.base = DLA(opt=opt)
        channels = self.base.channels
        scales = [2 ** i for i in range(len(channels[self.first_level:]))]
        self.dla_up = DLAUp(self.first_level, channels[self.first_level:], scales)
        self.ida_up = IDAUp(channels[self.first_level], channels[self.first_level:self.last_level],
                            [2 ** i for i in range(self.last_level - self.first_level)])
        self.K_list = [int(getattr(opt, "k_list_%d" % (i + 1))) for i in range(6)]
        self.kernel_list = [int(getattr(opt, "ks%d" % (i + 1))) for i in range(6)]
        nc = opt.num_classes
        self.transformer = nn.ModuleList([
            TransformerEncoder(TransformerEncoderLayer(
                d_inp=16 * 2 ** i, d_model=4 * 2 ** i,
                n_k=nc * self.K_list[i] * (1 + 2 * (self.kernel_list[i] // 2)) ** 2,
                pos_embed=getattr(opt, "pos_embed", True)), num_layers=3) for i in range(3)])
        self.cat_layer = nn.ModuleList([
            nn.Sequential(nn.Linear(16 * 2 ** (i + 1), 32 * 2 ** (i + 1)), nn.ReLU(),
                          nn.Linear(32 * 2 ** (i + 1), 16 * 2 ** i)) for i in range(6)])
        # inference-only switch: levels 0/1 never reach the output (DLAUp starts at level 2)
        self.skip_dead_levels = False

    def img2feats(self, x):
        raise NotImplementedError

    def fuse_level(self, i, pre_feats, cur_feats, pre_flat, rep_flat, Whm):
        B, C, H, W = cur_feats.shape
        pre_ids = fusion.window_ids(pre_flat[i], Whm, self.SCALE_LIST[i], self.kernel_list[i], H, W)
        cur_ids = fusion.window_ids(rep_flat[i], Whm, self.SCALE_LIST[i], self.kernel_list[i], H, W)
        pre_key = fusion.gather_tokens(pre_feats, pre_ids)
        cur_query = fusion.gather_tokens(cur_feats, cur_ids)
        out = self.transformer[i](cur_query, pre_key, pre_key) if i <= 2 else pre_key
        rows = self.cat_layer[i](torch.cat([out, cur_query], dim=-1))
        return fusion.scatter_tokens(cur_feats, cur_ids, rows), pre_ids, cur_ids

    def imgpre2feats(self, x, pre_img=None, pre_hm=None, repro_hm=None, pre_hm_cls=None,
                     repro_hm_cls=None):
        x_pre = self.base(pre_img=pre_img, pre_hm=pre_hm)
        x_cur = self.base(pre_img=x, pre_hm=repro_hm)
        Whm = pre_hm_cls.shape[3]
        pre_flat, rep_flat = {}, {}
        for K in set(self.K_list):
            p, r = fusion.topk_flat_index(pre_hm_cls, K), fusion.topk_flat_index(repro_hm_cls, K)
            for i in range(6):
                if self.K_list[i] == K:
                    pre_flat[i], rep_flat[i] = p, r
        x_out, pre_ids_0, cur_ids_0 = [], None, None
        for i in range(6):
            if self.skip_dead_levels and i < self.first_level and not self.training:
                x_out.append(x_cur[i])
                continue
            f, pid, cid = self.fuse_level(i, x_pre[i], x_cur[i], pre_flat, rep_flat, Whm)
            if i == 0:
                pre_ids_0, cur_ids_0 = pid, cid
            x_out.append(f)
        x_out = self.dla_up(x_out)
        y = [x_out[i].clone() for i in range(self.last_level - self.first_level)]
        self.ida_up(y, 0, len(y))
        return [y[-1]], pre_ids_0, cur_ids_0

    def forward(self, x, pre_img=None, pre_hm=None, repro_hm=None, pre_hm_cls=None, repro_hm_cls=None):
        # the plain convolutions of this module tree are cuDNN calls, which default to TF32 (1e-3 per op):
        # keep them in fp32 so the tree meets the same 1e-3 end-to-end bound as the compiled engine
        cd = torch.backends.cudnn
        with cd.flags(enabled=cd.enabled, benchmark=cd.benchmark, deterministic=cd.deterministic, allow_tf32=False):
            return self._forward(x, pre_img, pre_hm, repro_hm, pre_hm_cls, repro_hm_cls)

    def _forward(self, x, pre_img, pre_hm, repro_hm, pre_hm_cls, repro_hm_cls):
        if all(t is None for t in (pre_img, pre_hm, repro_hm, pre_hm_cls, repro_hm_cls)):
            feats = self.img2feats(x)
        else:
            feats, _, _ = self.imgpre2feats(x, pre_img, pre_hm, repro_hm, pre_hm_cls, repro_hm_cls)
        out = []
        for s in range(self.num_stacks):
            names = [h for h in self.heads if "wh" not in h]
            if getattr(self.opt, "model_output_list", False):
                out.append([getattr(self, h)(feats[s]) for h in sorted(names)])
            else:
                out.append({h: getattr(self, h)(feats[s]) for h in names})
        return out


_network_factory = {"dlapawdl3new": DLA_PlanAWindow_l3new}


def create_model(arch, head, head_conv, opt=None):
    """model.py:24-29."""
    num_layers = int(arch[arch.find("_") + 1:]) if "_" in arch else 0
    name = arch[:arch.find("_")] if "_" in arch else arch
    if name not in _network_factory:
        raise KeyError("arch %r is outside the built hot path (only dlapawdl3new_34)" % arch)
    return _network_factory[name](num_layers, heads=head, head_convs=head_conv, opt=opt)


def load_model(model, model_path, opt=None, optimizer=None):
    """Tolerant checkpoint load, same call forms and return values as model.py:43-103: strips a leading `module.`;
    on a shape mismatch (or `opt.reset_hm` for 80- / 1-class `hm*` tensors) either re-uses the leading slice
    (`opt.reuse_hm`) or keeps the model's own tensor; missing keys keep their init.  With an optimizer the result
    is `(model, optimizer, start_epoch)` and, under `opt.resume`, the learning rate is stepped down by 0.1 for
    every `opt.lr_step` already passed (the reference does not restore the optimizer state either, :86)."""
    start_epoch = 0
    ckpt = torch.load(model_path, map_location="cpu")
    src = ckpt.get("state_dict", ckpt)
    src = {(k[7:] if k.startswith("module") and not k.startswith("module_list") else k): v
           for k, v in src.items()}
    own = model.state_dict()
    reset_hm, reuse_hm = bool(getattr(opt, "reset_hm", False)), bool(getattr(opt, "reuse_hm", False))
    merged = {}
    for k, v in src.items():
        if k not in own:
            print("Drop parameter {}.".format(k))
            continue
        if v.shape != own[k].shape or (reset_hm and k.startswith("hm") and v.shape[0] in (80, 1)):
            if reuse_hm and v.dim() == own[k].dim() and v.shape[1:] == own[k].shape[1:]:
                print("Reusing parameter {}, required shape{}, loaded shape{}.".format(k, own[k].shape, v.shape))
                t = own[k].clone()
                n = min(t.shape[0], v.shape[0])
                t[:n] = v[:n]
                merged[k] = t
            else:
                print("Skip loading parameter {}, required shape{}, loaded shape{}.".format(k, own[k].shape, v.shape))
                merged[k] = own[k]
        else:
            merged[k] = v
    for k, v in own.items():
        if k not in merged:
            print("No param {}.".format(k))
            merged[k] = v
    model.load_state_dict(merged, strict=False)
    if optimizer is not None and getattr(opt, "resume", False):
        if "optimizer" in ckpt:
            start_epoch = ckpt["epoch"]
            start_lr = opt.lr
            for step in opt.lr_step:
                if start_epoch >= step:
                    start_lr *= 0.1
            for group in optimizer.param_groups:
                group["lr"] = start_lr
            print("Resumed optimizer with start lr", start_lr)
        else:
            print("No optimizer parameters in checkpoint.")
    if optimizer is not None:
        return model, optimizer, start_epoch
    return model


def save_model(path, epoch, model, optimizer=None):
    """model.py:105-114."""
    wrapped = isinstance(model, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel))
    data = {"epoch": epoch, "state_dict": (model.module if wrapped else model).state_dict()}
    if optimizer is not None:
        data["optimizer"] = optimizer.state_dict()
    torch.save(data, path)
