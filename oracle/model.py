"""Oracle: the whole network forward, restated functionally over a reference state-dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  fp32, CPU, torch ops.

Restates ``BaseModelPlanA.forward`` (reference
``sgtapose/lib/model/networks/base_model.py:170-200``) ->
``DLA_PlanAWindow_l3new.imgpre2feats`` (``dla.py:1505-1554``) for the shipped
configuration (``--arch dlapawdl3new_34 --dla_node dcn``, eval mode), consuming
the reference's own parameter names, so one state-dict drives the reference
model, this oracle and the CUDA engine.  Pinned against the imported reference
by tests/test_oracle_model.py and tests/golden/model_*.npz.

Defined behaviours where the reference is nondeterministic / tie-dependent
(SURVEY.md 7.3 H3, H5): top-k ties -> value descending, lowest index first;
duplicate scatter indices in ``substitute_topk_features_scale`` (dla.py:1014)
-> the HIGHEST token index wins (== sequential CPU ``index_put_``).
"""
import math

import torch
import torch.nn.functional as F

from . import dcn as odcn

CHANNELS = [16, 32, 64, 128, 256, 512]            # dla.py:355-356
SCALE_LIST = [4, 2, 1, 1 / 2, 1 / 4, 1 / 8]       # dla.py:1493
N_HEADS = 8                                       # dla.py:705
BN_EPS = 1e-5


# ------------------------------------------------------------------ primitives
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def _conv(sd, p, x, stride=1, pad=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=pad)


def _conv_bn_relu(sd, pc, pb, x, stride=1, pad=1):
    return F.relu(_bn(sd, pb, _conv(sd, pc, x, stride, pad)))


def _basic_block(sd, p, x, residual, stride):
    """BasicBlock.forward, dla.py:55-69."""
    out = _conv_bn_relu(sd, p + ".conv1", p + ".bn1", x, stride, 1)
    out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out, 1, 1))
    return F.relu(out + residual)


def _root(sd, p, xs):
    """Root.forward, dla.py:167-175 (root_residual is False for dla34)."""
    return F.relu(_bn(sd, p + ".bn", _conv(sd, p + ".conv", torch.cat(xs, 1))))


def _tree(sd, p, x, levels, cin, cout, stride, level_root, children=None):
    """Tree.forward, dla.py:218-231."""
    children = [] if children is None else children
    bottom = F.max_pool2d(x, stride, stride) if stride > 1 else x
    if levels == 1:
        residual = _bn(sd, p + ".project.1", _conv(sd, p + ".project.0", bottom)) if cin != cout else bottom
        if level_root:
            children.append(bottom)
        x1 = _basic_block(sd, p + ".tree1", x, residual, stride)
        x2 = _basic_block(sd, p + ".tree2", x1, x1, 1)
        return _root(sd, p + ".root", [x2, x1] + children)
    # levels == 2: the projected residual is computed by the reference but never used
    if level_root:
        children.append(bottom)
    x1 = _tree(sd, p + ".tree1", x, levels - 1, cin, cout, stride, False)
    children.append(x1)
    return _tree(sd, p + ".tree2", x1, levels - 1, cout, cout, 1, False, children)


def dla34_base(sd, img, hm, p="base"):
    """DLA.forward with x=None (dla.py:324-337): pre_img_layer(img) + pre_hm_layer(hm),
    then level0..5.  ``base_layer`` is never executed on this path (dla.py:1506-1507)."""
    x = _conv_bn_relu(sd, p + ".pre_img_layer.0", p + ".pre_img_layer.1", img, 1, 3) \
        + _conv_bn_relu(sd, p + ".pre_hm_layer.0", p + ".pre_hm_layer.1", hm, 1, 3)
    ys = []
    x = _conv_bn_relu(sd, p + ".level0.0", p + ".level0.1", x, 1, 1); ys.append(x)
    x = _conv_bn_relu(sd, p + ".level1.0", p + ".level1.1", x, 2, 1); ys.append(x)
    x = _tree(sd, p + ".level2", x, 1, 32, 64, 2, False); ys.append(x)
    x = _tree(sd, p + ".level3", x, 2, 64, 128, 2, True); ys.append(x)
    x = _tree(sd, p + ".level4", x, 2, 128, 256, 2, True); ys.append(x)
    x = _tree(sd, p + ".level5", x, 1, 256, 512, 2, True); ys.append(x)
    return ys


def deform_conv_block(sd, p, x, use_torchvision=True):
    """DeformConv.forward, dla.py:547-550: DCN -> BN -> ReLU."""
    f = odcn.dcn_forward_torchvision if use_torchvision else odcn.dcn_forward
    y = f(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"],
          sd[p + ".conv.conv_offset_mask.weight"], sd[p + ".conv.conv_offset_mask.bias"])
    return F.relu(_bn(sd, p + ".actf.0", y))


def _up(sd, p, x):
    """depth-wise ConvTranspose2d(o, o, 2f, stride=f, padding=f//2, groups=o), dla.py:561-563."""
    w = sd[p + ".weight"]
    f = w.shape[2] // 2
    return F.conv_transpose2d(x, w, None, stride=f, padding=f // 2, groups=w.shape[0])


def ida_up(sd, p, layers, startp, endp, tv=True):
    """IDAUp.forward, dla.py:571-577 (mutates ``layers``)."""
    for i in range(startp + 1, endp):
        k = str(i - startp)
        layers[i] = _up(sd, p + ".up_" + k, deform_conv_block(sd, p + ".proj_" + k, layers[i], tv))
        layers[i] = deform_conv_block(sd, p + ".node_" + k, layers[i] + layers[i - 1], tv)


def dla_up(sd, p, layers, tv=True):
    """DLAUp.forward, dla.py:600-606 with startp=2."""
    out = [layers[-1]]
    for i in range(len(layers) - 2 - 1):
        ida_up(sd, "%s.ida_%d" % (p, i), layers, len(layers) - i - 2, len(layers), tv)
        out.insert(0, layers[-1])
    return out


# ------------------------------------------------------------------ attention path
def topk_index(hm_cls, K):
    """get_topk_index, dla.py:898-913: per (sample, channel) top-K flat index ->
    (x = id % W, y = id // W) as fp32 [B, C*K, 2].  Ties: lowest index first."""
    B, C, H, W = hm_cls.shape
    flat = hm_cls.reshape(B, C, H * W)
    order = torch.sort(flat, dim=-1, descending=True, stable=True)[1]
    idx = order[..., :K]
    idx = idx.reshape(B, C * K)
    return torch.stack([(idx % W).float(), (idx // W).float()], -1)


def window_ids(topk_inds, scale, kernel, H, W):
    """get_topk_features_scale's index arithmetic, dla.py:932-957, INCLUDING the fp32
    detour: coords = ind*scale + offset (fp32), clamp [0,H-1],
    id = y*W + x in fp32, then truncation to int64.  Offsets: x-major meshgrid."""
    r = torch.arange(-(kernel // 2), kernel // 2 + 1, dtype=torch.float32)
    n = len(r)
    off = torch.stack([r[:, None].expand(n, n).reshape(-1), r[None, :].expand(n, n).reshape(-1)], -1)
    coords = topk_inds[:, :, None, :] * scale + off[None, None]
    coords = coords.clamp(0, H - 1)
    fid = coords[..., 1] * W + coords[..., 0]
    return fid.reshape(topk_inds.shape[0], -1).long()        # [B, C*K*n*n]


def gather_tokens(feats, fid):
    """feats [B,C,H,W] -> rows of the NHWC view at fid: [B, n, C] (dla.py:961-966)."""
    B, C, H, W = feats.shape
    flat = feats.reshape(B, C, H * W).permute(0, 2, 1)
    return torch.gather(flat, 1, fid[..., None].expand(-1, -1, C))


def mhca(sd, p, q_in, kv_in, use_pos=True):
    """MHCA_ein.forward, dla.py:868-887."""
    Q = F.linear(q_in, sd[p + ".w_q.weight"])
    K = F.linear(kv_in, sd[p + ".w_k.weight"])
    V = F.linear(kv_in, sd[p + ".w_v.weight"])
    B, n, hid = Q.shape
    d = hid // N_HEADS
    Q = Q.reshape(B, n, N_HEADS, d).permute(0, 2, 1, 3)
    K = K.reshape(B, n, N_HEADS, d).permute(0, 2, 1, 3)
    V = V.reshape(B, n, N_HEADS, d).permute(0, 2, 1, 3)
    energy = torch.matmul(Q, K.transpose(-1, -2)) / math.sqrt(d)
    if use_pos:
        energy = energy + sd[p + ".pos_embed"]
    attn = torch.softmax(energy, -1)
    out = torch.matmul(attn, V).permute(0, 2, 1, 3).reshape(B, n, hid)
    return F.linear(out, sd[p + ".fc.weight"], sd[p + ".fc.bias"])


def encoder_layer(sd, p, query, key, use_pos=True):
    """TransformerEncoderLayer.forward (eval: dropout = identity), dla.py:728-743."""
    C = query.shape[-1]
    q = mhca(sd, p + ".cross_attn", query, key, use_pos) + query
    q = F.layer_norm(q, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    f = F.linear(F.relu(F.linear(q, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                 sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    return F.layer_norm(q + f, (C,), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])


def cat_mlp(sd, p, a, b):
    h = F.relu(F.linear(torch.cat([a, b], -1), sd[p + ".0.weight"], sd[p + ".0.bias"]))
    return F.linear(h, sd[p + ".2.weight"], sd[p + ".2.bias"])


def scatter_tokens(feats, fid, rows):
    """substitute_topk_features_scale's write-back (dla.py:1014-1016) with the
    defined duplicate rule: highest token index wins (sequential assignment)."""
    B, C, H, W = feats.shape
    flat = feats.reshape(B, C, H * W).permute(0, 2, 1).contiguous()
    for b in range(B):
        ids = fid[b].tolist()
        last = {}
        for t, i in enumerate(ids):
            last[i] = t
        tgt = torch.tensor(list(last.keys()), dtype=torch.long)
        src = torch.tensor(list(last.values()), dtype=torch.long)
        flat[b, tgt] = rows[b, src]
    return flat.permute(0, 2, 1).reshape(B, C, H, W).contiguous()


def fuse_level(sd, i, pre_feats, cur_feats, pre_hm_cls, repro_hm_cls, K_list, kernel_list,
               use_pos=True, n_layers=3):
    """One iteration of the loop at dla.py:1513-1543."""
    B, C, H, W = cur_feats.shape
    pre_idx = topk_index(pre_hm_cls, K_list[i])
    rep_idx = topk_index(repro_hm_cls, K_list[i])
    pre_fid = window_ids(pre_idx, SCALE_LIST[i], kernel_list[i], H, W)
    cur_fid = window_ids(rep_idx, SCALE_LIST[i], kernel_list[i], H, W)
    pre_key = gather_tokens(pre_feats, pre_fid)
    cur_query = gather_tokens(cur_feats, cur_fid)
    if i <= 2:
        out = cur_query
        for _ in range(n_layers):                       # 3 layers, ONE shared module (dla.py:788-789)
            out = encoder_layer(sd, "transformer.%d.layers.0" % i, out, pre_key, use_pos)
    else:
        out = pre_key
    rows = cat_mlp(sd, "cat_layer.%d" % i, out, cur_query)
    return scatter_tokens(cur_feats, cur_fid, rows), pre_fid, cur_fid


# ------------------------------------------------------------------ whole forward
def head(sd, name, x):
    """base_model.py:121-135: conv3x3 64->256 + ReLU + conv1x1."""
    return _conv(sd, name + ".2", F.relu(_conv(sd, name + ".0", x, 1, 1)))


def forward(sd, x, pre_img, pre_hm, repro_hm, pre_hm_cls, repro_hm_cls,
            K_list=(1,) * 6, kernel_list=(12, 6, 3, 1, 1, 1), use_pos=True,
            heads=("hm", "reg", "tracking"), tv=True, return_feats=False, grad=False):
    """-> [ {hm, reg, tracking} ] exactly like BaseModelPlanA.forward (no sigmoid).
    grad=True keeps the autograd graph (training-step parity: the oracle for backward is torch autograd
    through these same ops, like the reference's Trainer, trainer_parallel.py:267-284)."""
    with torch.set_grad_enabled(grad):
        x_pre = dla34_base(sd, pre_img, pre_hm)
        x_cur = dla34_base(sd, x, repro_hm)
        fused = []
        for i in range(6):
            f, _, _ = fuse_level(sd, i, x_pre[i], x_cur[i], pre_hm_cls, repro_hm_cls,
                                 K_list, kernel_list, use_pos)
            fused.append(f)
        outs = dla_up(sd, "dla_up", list(fused), tv)
        y = [outs[0].clone(), outs[1].clone(), outs[2].clone()]
        ida_up(sd, "ida_up", y, 0, 3, tv)
        feat = y[-1]
        z = {h: head(sd, h, feat) for h in heads}
    if return_feats:
        return [z], {"x_pre": x_pre, "x_cur": x_cur, "fused": fused, "feat": feat}
    return [z]
