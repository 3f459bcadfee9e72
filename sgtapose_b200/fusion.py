"""Structure-prior temporal attention: host side of the token path and attention core.

Mirrors, with the same names and argument meaning, the reference functions in
sgtapose/lib/model/networks/dla.py: get_topk_index (:898-913), get_topk_features_scale
(:915-968), substitute_topk_features_scale (:1006-1018) and MHCA_ein (:848-887), but every
step runs in a CUDA kernel of libsgta_b200.so; no index tensor ever visits the host.

Defined behaviour where the reference is tie-/order-dependent:
top-k ties -> lowest index first; duplicate write-back indices -> highest token index wins.
"""
import math

import torch
from torch import nn

from . import _lib


def topk_flat_index(hm, K):
    """[B,C,H,W] fp32 -> [B, C*K] int64 flat indices (value desc, index asc)."""
    B, C, H, W = hm.shape
    hm = hm.contiguous().float()
    idx = torch.empty(B, C * K, device=hm.device, dtype=torch.int64)
    _lib.call("sgta_topk_index", _lib.ptr(hm), _lib.ptr(idx), B, C, H * W, K, _lib.stream())
    return idx


def get_topk_index(pre_hm, repro_hm, K):
    """dla.py:898-913.  Returns ([B,C*K,2], [B,C*K,2]) fp32 (x, y) like the reference, but as
    CUDA tensors (the reference returns CPU tensors and pays a sync per level)."""
    assert pre_hm.shape == repro_hm.shape
    W = pre_hm.shape[3]
    out = []
    for hm in (pre_hm, repro_hm):
        idx = topk_flat_index(hm, K)
        out.append(torch.stack([(idx % W).float(), torch.div(idx, W, rounding_mode="floor").float()], -1))
    return out[0], out[1]


def window_ids(flat_idx, Whm, scale, kernel, H, W):
    """[B,CK] int64 flat prior indices -> [B, CK*win^2] int64 feature-map ids with the
    reference's fp32 index arithmetic (dla.py:932-957)."""
    B, CK = flat_idx.shape
    win = 2 * (kernel // 2) + 1
    ids = torch.empty(B, CK * win * win, device=flat_idx.device, dtype=torch.int64)
    _lib.call("sgta_window_ids", _lib.ptr(flat_idx), _lib.ptr(ids), B, CK, Whm, float(scale),
              kernel, H, W, _lib.stream())
    return ids


def gather_tokens(feats, ids):
    """rows[b,t,:] = feats[b,:,ids[b,t]]; feats NCHW fp32.  Differentiable w.r.t. feats."""
    if torch.is_grad_enabled() and feats.requires_grad:
        B, C, H, W = feats.shape
        flat = feats.reshape(B, C, H * W).permute(0, 2, 1)
        return torch.gather(flat, 1, ids[..., None].expand(-1, -1, C))
    B, C, H, W = feats.shape
    feats = feats.contiguous().float()
    n = ids.shape[1]
    rows = torch.empty(B, n, C, device=feats.device, dtype=torch.float32)
    _lib.call("sgta_gather_tokens", _lib.ptr(feats), _lib.ptr(ids), _lib.ptr(rows), B, C, H * W, n, 0,
              _lib.stream())
    return rows


def _last_writer_mask(ids):
    """mask[b,t] = no later token carries the same id (training-time twin of the kernel rule)."""
    B, n = ids.shape
    order = torch.arange(n, device=ids.device)
    key = ids * n + order[None]
    srt, perm = torch.sort(key, dim=1)
    sid = torch.div(srt, n, rounding_mode="floor")
    last = torch.ones_like(sid, dtype=torch.bool)
    last[:, :-1] = sid[:, :-1] != sid[:, 1:]
    mask = torch.zeros_like(last)
    mask.scatter_(1, perm, last)
    return mask


def scatter_tokens(feats, ids, rows):
    """Returns a copy of feats with feats[b,:,ids[b,t]] = rows[b,t,:]; highest t wins."""
    B, C, H, W = feats.shape
    if torch.is_grad_enabled() and (feats.requires_grad or rows.requires_grad):
        mask = _last_writer_mask(ids)
        flat = feats.reshape(B, C, H * W).permute(0, 2, 1)
        bidx = torch.arange(B, device=ids.device)[:, None].expand_as(ids)
        out = flat.index_put((bidx[mask], ids[mask]), rows[mask])
        return out.permute(0, 2, 1).reshape(B, C, H, W).contiguous()
    out = feats.contiguous().float().clone()
    rows = rows.contiguous().float()
    n = ids.shape[1]
    _lib.call("sgta_scatter_tokens", _lib.ptr(out), _lib.ptr(ids), _lib.ptr(rows), B, C, H * W, n, 0,
              _lib.stream())
    return out


def get_topk_features_scale(feats, topk_inds, scale_num, kernel=3):
    """dla.py:915-968.  topk_inds: [B,K,2] fp32 (x,y) as returned by get_topk_index.
    Returns (selected_feat [B,K*N,C], batch_id [B,K*N], feat_id [B,K*N])."""
    B, C, H, W = feats.shape
    assert H == W
    flat = (topk_inds[..., 1].long() * 65536 + topk_inds[..., 0].long()).contiguous()
    ids = window_ids(flat, 65536, scale_num, kernel, H, W)
    batch_id = torch.arange(B, device=feats.device)[:, None].expand_as(ids)
    return gather_tokens(feats, ids), batch_id, ids


def substitute_topk_features_scale(out, cur_features, batch_id, feat_id, mlp):
    """dla.py:1006-1018."""
    cur_query = gather_tokens(cur_features, feat_id)
    rows = mlp(torch.cat([out, cur_query], dim=-1))
    return scatter_tokens(cur_features, feat_id, rows)


class _AttnCore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, pos, heads, inv_scale):
        if not q.is_cuda:
            raise _lib.SgtaError("attention: inputs must be CUDA tensors (no CPU fallback)")
        q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        pos_c = pos.contiguous().float() if pos is not None else None
        B, nq, HD = q.shape
        nk = k.shape[1]
        d = HD // heads
        out = torch.empty_like(q)
        _lib.call("sgta_attn_forward", _lib.ptr(q), _lib.ptr(k), _lib.ptr(v), _lib.ptr(pos_c),
                  _lib.ptr(out), B, heads, nq, nk, d, float(inv_scale), _lib.stream())
        ctx.save_for_backward(q, k, v, pos_c if pos_c is not None else q.new_empty(0))
        ctx.cfg = (heads, inv_scale, pos is not None)
        return out

    @staticmethod
    def backward(ctx, go):
        q, k, v, pos = ctx.saved_tensors
        heads, inv_scale, has_pos = ctx.cfg
        go = go.contiguous().float()
        B, nq, HD = q.shape
        nk = k.shape[1]
        gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        gpos = torch.zeros_like(pos) if (has_pos and ctx.needs_input_grad[3]) else None
        _lib.call("sgta_attn_backward", _lib.ptr(q), _lib.ptr(k), _lib.ptr(v),
                  _lib.ptr(pos) if has_pos else None, _lib.ptr(go), _lib.ptr(gq), _lib.ptr(gk),
                  _lib.ptr(gv), _lib.ptr(gpos), B, heads, nq, nk, HD // heads, float(inv_scale),
                  _lib.stream())
        return gq, gk, gv, gpos, None, None


def attention_core(q, k, v, pos, heads, scale):
    """softmax(q k^T / scale + pos) v on "b n (h d)" tensors, fused on the device."""
    return _AttnCore.apply(q, k, v, pos, heads, 1.0 / scale)


class MHCA_ein(nn.Module):
    """dla.py:848-887, same parameters (w_q, w_k, w_v, fc, pos_embed)."""

    def __init__(self, num_heads, inp_dim, hid_dim, n, pos_embed=True):
        super().__init__()
        assert hid_dim % num_heads == 0
        self.hid_dim, self.inp_dim, self.n_heads, self.n = hid_dim, inp_dim, num_heads, n
        self.pos_embed_bool = pos_embed
        self.w_q = nn.Linear(inp_dim, hid_dim, bias=False)
        self.w_k = nn.Linear(inp_dim, hid_dim, bias=False)
        self.w_v = nn.Linear(inp_dim, hid_dim, bias=False)
        self.fc = nn.Linear(hid_dim, inp_dim)
        self.scale = math.sqrt(hid_dim // num_heads)
        self.pos_embed = nn.Parameter(torch.zeros(num_heads, n, n))

    def forward(self, query, key, value):
        pos = self.pos_embed if (self.pos_embed is not None and self.pos_embed_bool) else None
        out = attention_core(self.w_q(query), self.w_k(key), self.w_v(value), pos, self.n_heads,
                             self.scale)
        return self.fc(out)


def token_linear(x, w, bias=None, x2=None, relu=False):
    """nn.Linear on token rows in our own kernel: act([x | x2] w^T + bias).  x [..., K1] (and x2 [..., K2]) fp32 CUDA,
    w [N, K1+K2] as nn.Linear stores it.  The K/V/first-Q projections (dla.py:868-876) and both layers of
    cat_layer (dla.py:1499-1502; the torch.cat of :1015 is folded into the first one).  Inference only."""
    x = x.contiguous()
    K1 = x.shape[-1]
    K2 = 0 if x2 is None else x2.shape[-1]
    if x2 is not None:
        x2 = x2.contiguous()
    N = w.shape[0]
    if w.shape[1] != K1 + K2:
        raise _lib.SgtaError("token_linear: weight [%d,%d] does not match K = %d + %d" % (N, w.shape[1], K1, K2))
    y = torch.empty(x.shape[:-1] + (N,), device=x.device, dtype=torch.float32)
    M = x.numel() // K1
    _lib.call("sgta_token_linear", _lib.ptr(x), K1, _lib.ptr(x2), K2, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), M, N,
              int(bool(relu)), _lib.stream())
    return y


def kv_head_major_supported(B, heads, nq, nk, d, has_pos=True):
    """Does the head-major K/V form of the attention core serve this shape (level 0 at the bench batch)?"""
    return bool(_lib.load().sgta_attn_kvhm_supported(B, heads, nq, nk, d, int(bool(has_pos))))


def token_linear_heads(x, w, heads):
    """w_k / w_v projection with a head-major result: x [B,n,K], w [N,K] -> [B, heads, n, N/heads] (inference only)."""
    x = x.contiguous()
    B, n, K = x.shape
    N = w.shape[0]
    y = torch.empty(B, heads, n, N // heads, device=x.device, dtype=torch.float32)
    _lib.call("sgta_token_linear_heads", _lib.ptr(x), K, _lib.ptr(w), _lib.ptr(y), B * n, N, n, heads, _lib.stream())
    return y


def attention_core_kvhm(q, k_hm, v_hm, pos, heads, scale):
    """attention_core with head-major K / V ([B, heads, nk, d], `token_linear_heads`); q and the result stay
    "b n (h d)".  Inference only; shapes as accepted by `kv_head_major_supported`."""
    q = q.contiguous()
    B, nq, HD = q.shape
    nk, d = k_hm.shape[2], k_hm.shape[3]
    out = torch.empty_like(q)
    _lib.call("sgta_attn_forward_kvhm", _lib.ptr(q), _lib.ptr(k_hm), _lib.ptr(v_hm), _lib.ptr(pos.contiguous()),
              _lib.ptr(out), B, heads, nq, nk, d, float(1.0 / scale), _lib.stream())
    return out


def token_mlp(att, q, fc_wt, fc_b, ln1_w, ln1_b, w1, b1, w2t, b2, ln3_w, ln3_b, wq_next=None, eps=1e-5):
    """Post-attention half of TransformerEncoderLayer.forward (dla.py:734-743) in ONE launch:
    q2 = LN3(q1 + FFN(q1)), q1 = LN1(fc(att) + q); also the next layer's query projection w_q q2 when
    `wq_next` is given.  att [B,n,hid], q [B,n,C] fp32 CUDA -> (q2 [B,n,C], qp [B,n,hid] | None).
    fc_wt = fc.weight.t().contiguous() [hid,C], w2t = linear2.weight.t().contiguous() [dffn,C] (made once).
    Inference only (no autograd): the module tree keeps the differentiable library path."""
    B, n, C = q.shape
    hid = att.shape[-1]
    att, q = att.contiguous(), q.contiguous()
    q_out = torch.empty_like(q)
    qp = torch.empty(B, n, hid, device=q.device, dtype=torch.float32) if wq_next is not None else None
    _lib.call("sgta_token_mlp", _lib.ptr(att), _lib.ptr(q), _lib.ptr(fc_wt), _lib.ptr(fc_b), _lib.ptr(ln1_w),
              _lib.ptr(ln1_b), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2t), _lib.ptr(b2), _lib.ptr(ln3_w),
              _lib.ptr(ln3_b), _lib.ptr(wq_next), _lib.ptr(q_out), _lib.ptr(qp), B * n, C, hid, w1.shape[0],
              float(eps), _lib.stream())
    return q_out, qp
