"""Compiled inference engine: the whole per-frame forward as a static plan of hand-written
sm_100a kernels over NHWC buffers, optionally captured in one CUDA graph.

Same parameters, same math and same outputs as the module tree in networks.py (which mirrors
reference sgtapose/lib/model/networks/dla.py:1505-1554 + base_model.py:170-200); what changes
is the execution:

  * eval-mode BatchNorm, conv bias and ReLU are folded into the epilogue of the producing
    tcgen05 implicit-GEMM kernel (conv_umma.cu); residual adds too;
  * the previous-frame and current-frame passes of the shared DLA-34 base (dla.py:1506-1507)
    run as ONE batch of 2B images; the two 7x7 stems (image + heat-map, dla.py:325-331) are
    one GEMM with a dual-ReLU epilogue;
  * Root concatenations (dla.py:169) are zero-copy: producers write channel slices of the
    concat buffer (pixel-stride arguments of the C ABI);
  * every DeformConv = offset/mask conv (N=27 padded to 32, fp32 out) + the fused
    bilinear-gather DCN GEMM; the depth-wise up-sampler and the IDAUp skip add are one kernel;
  * token selection, gather, attention core, write-back and decode never leave the device;
  * the three head convs 64->256 run as one 64->768 GEMM; the 1x1 output convs write NCHW fp32
    directly (optionally with the detector's sigmoid, sgta_detector.py:854-862, fused).

mode="fp32": fp32 activations, F32X3 split-bf16 MMAs (parity bound 1e-3 relative);
mode="bf16": bf16 activations and MMAs, fp32 accumulate (stated looser bound).
"""
import torch
import torch.nn.functional as F

from . import _lib, decode, fastops as fo, fusion

SCALE_LIST = (4, 2, 1, 1 / 2, 1 / 4, 1 / 8)
BN_EPS = 1e-5


def _fold_bn(sd, p, bias=None):
    g, b = sd[p + ".weight"].float(), sd[p + ".bias"].float()
    m, v = sd[p + ".running_mean"].float(), sd[p + ".running_var"].float()
    scale = g / torch.sqrt(v + BN_EPS)
    shift = b - m * scale
    if bias is not None:
        shift = shift + bias.float() * scale
    return scale, shift


class InferenceEngine:
    def __init__(self, state_dict, opt, batch, size=384, mode="fp32", device="cuda", fuse_sigmoid=False,
                 skip_dead_levels=False, use_graph=True):
        if mode not in ("fp32", "bf16"):
            raise ValueError("mode must be 'fp32' or 'bf16'")
        if size % 32:
            raise ValueError("input size must be a multiple of 32")
        self.dev = torch.device(device)
        self.B, self.S, self.mode = batch, size, mode
        self.mma = fo.MMA_F32X3 if mode == "fp32" else fo.MMA_BF16
        self.adt = torch.float32 if mode == "fp32" else torch.bfloat16
        self.opt = opt
        self.fuse_sigmoid = fuse_sigmoid
        self.skip_dead_levels = skip_dead_levels
        self.K_list = [int(getattr(opt, "k_list_%d" % (i + 1))) for i in range(6)]
        self.kernel_list = [int(getattr(opt, "ks%d" % (i + 1))) for i in range(6)]
        self.use_pos = bool(getattr(opt, "pos_embed", True))
        self.sd = {k: v.detach().to(self.dev) for k, v in state_dict.items()}
        self._build_specs()
        self._alloc()
        self.graph = None
        self._use_graph = use_graph
        self._warm = False

    # ------------------------------------------------------------------ weights
    def _conv_bn(self, pc, pb, stride, pad, act):
        w = self.sd[pc + ".weight"].float()
        Cout, Cin, kh, kw = w.shape
        scale, shift = _fold_bn(self.sd, pb)
        return fo.ConvSpec(fo.weight_matrix(w), scale, shift, Cin, kh, kw, stride, pad, self.mma, act)

    def _block(self, p, stride):
        return {"c1": self._conv_bn(p + ".conv1", p + ".bn1", stride, 1, fo.ACT_RELU),
                "c2": self._conv_bn(p + ".conv2", p + ".bn2", 1, 1, fo.ACT_RELU)}   # relu after +residual

    def _deform(self, p):
        sd = self.sd
        w = sd[p + ".conv.weight"].float()
        Cout, Cin = w.shape[:2]
        omw, omb = sd[p + ".conv.conv_offset_mask.weight"].float(), sd[p + ".conv.conv_offset_mask.bias"].float()
        om = fo.ConvSpec(fo.weight_matrix(omw), torch.ones(27, device=self.dev), omb, Cin, 3, 3, 1, 1, self.mma,
                         fo.ACT_NONE)                                # N padded 27 -> 32
        scale, shift = _fold_bn(sd, p + ".actf.0", sd[p + ".conv.bias"])
        return {"om": om, "w": fo.pack_dcn_weight(w, self.mma), "scale": scale.contiguous(),
                "shift": shift.contiguous(), "Cin": Cin, "Cout": Cout}

    def _build_specs(self):
        sd, mma = self.sd, self.mma
        s = {}
        # stem: one GEMM over [img(3) | hm(1)] channels, N = 16 (image conv) + 16 (heat-map conv)
        wi, wh = sd["base.pre_img_layer.0.weight"].float(), sd["base.pre_hm_layer.0.weight"].float()
        wm = torch.cat([fo.weight_matrix(wi, 4, 0), fo.weight_matrix(wh, 4, 3)], 0)
        sa, ta = _fold_bn(sd, "base.pre_img_layer.1")
        sb, tb = _fold_bn(sd, "base.pre_hm_layer.1")
        s["stem"] = fo.ConvSpec(wm, torch.cat([sa, sb]), torch.cat([ta, tb]), 4, 7, 7, 1, 3, mma)
        s["level0"] = self._conv_bn("base.level0.0", "base.level0.1", 1, 1, fo.ACT_RELU)
        s["level1"] = self._conv_bn("base.level1.0", "base.level1.1", 2, 1, fo.ACT_RELU)

        def tree1(p, stride, project):
            t = {"tree1": self._block(p + ".tree1", stride), "tree2": self._block(p + ".tree2", 1),
                 "root": self._conv_bn(p + ".root.conv", p + ".root.bn", 1, 0, fo.ACT_RELU)}
            if project:
                t["project"] = self._conv_bn(p + ".project.0", p + ".project.1", 1, 0, fo.ACT_NONE)
            return t

        s["level2"] = tree1("base.level2", 2, True)
        for lv in (3, 4):
            p = "base.level%d" % lv
            s["level%d" % lv] = {"tree1": tree1(p + ".tree1", 2, True), "tree2": tree1(p + ".tree2", 1, False)}
        s["level5"] = tree1("base.level5", 2, True)
        # DLAUp / IDAUp
        for name, n_nodes in (("dla_up.ida_0", 1), ("dla_up.ida_1", 2), ("dla_up.ida_2", 3), ("ida_up", 2)):
            for k in range(1, n_nodes + 1):
                s["%s.proj_%d" % (name, k)] = self._deform("%s.proj_%d" % (name, k))
                s["%s.node_%d" % (name, k)] = self._deform("%s.node_%d" % (name, k))
                s["%s.up_%d" % (name, k)] = sd["%s.up_%d.weight" % (name, k)].float().contiguous()
        # heads: 64 -> 3 x 256 in one GEMM, then three 1x1 convs writing NCHW fp32
        self.head_names = [h for h in ("hm", "reg", "tracking") if (h + ".0.weight") in sd]
        w0 = torch.cat([fo.weight_matrix(sd[h + ".0.weight"].float()) for h in self.head_names], 0)
        b0 = torch.cat([sd[h + ".0.bias"].float() for h in self.head_names])
        s["head0"] = fo.ConvSpec(w0, torch.ones_like(b0), b0, 64, 3, 3, 1, 1, mma, fo.ACT_RELU)
        for h in self.head_names:
            w2, b2 = sd[h + ".2.weight"].float(), sd[h + ".2.bias"].float()
            act = fo.ACT_SIGMOID if (h == "hm" and self.fuse_sigmoid) else fo.ACT_NONE
            s["head2." + h] = fo.ConvSpec(fo.weight_matrix(w2), torch.ones_like(b2), b2, 256, 1, 1, 1, 0, mma,
                                          act, n_valid=w2.shape[0])
        self.specs = s

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        B, S, dev, adt = self.B, self.S, self.dev, self.adt
        B2 = 2 * B
        z = lambda *shape, dtype=adt: torch.zeros(*shape, device=dev, dtype=dtype)
        b = {}
        b["in4"] = z(B2, S, S, 4)
        b["f0"] = z(B2, S, S, 16)
        b["l0"] = z(B2, S, S, 16)
        b["l1"] = z(B2, S // 2, S // 2, 32)
        h2, h3, h4, h5 = S // 4, S // 8, S // 16, S // 32
        b["bot2"] = z(B2, h2, h2, 32)
        b["res2"] = z(B2, h2, h2, 64)
        b["mid2"] = z(B2, h2, h2, 64)
        b["cat2"] = z(B2, h2, h2, 128)
        b["l2"] = z(B2, h2, h2, 64)
        for lv, c, h in ((3, 128, h3), (4, 256, h4)):
            b["res%da" % lv] = z(B2, h, h, c)
            b["mid%d" % lv] = z(B2, h, h, c)
            b["cat%da" % lv] = z(B2, h, h, 2 * c)
            b["cat%db" % lv] = z(B2, h, h, 3 * c + c // 2)
            b["l%d" % lv] = z(B2, h, h, c)
        b["res5"] = z(B2, h5, h5, 512)
        b["mid5"] = z(B2, h5, h5, 512)
        b["cat5"] = z(B2, h5, h5, 1280)
        b["l5"] = z(B2, h5, h5, 512)
        for i, (c, h) in enumerate(((16, S), (32, S // 2), (64, h2), (128, h3), (256, h4), (512, h5))):
            b["fused%d" % i] = z(B, h, h, c)
        b["om"] = z(B, h2, h2, 32, dtype=torch.float32)          # largest DCN map; reused by all 16
        # DLAUp / IDAUp intermediates: proj output (at source res), sum (after up + skip), node output
        def trio(name, c, hs, f):
            b[name + ".p"] = z(B, hs, hs, c)
            b[name + ".s"] = z(B, hs * f, hs * f, c)
            b[name + ".n"] = z(B, hs * f, hs * f, c)
        trio("dla_up.ida_0.1", 256, h5, 2)
        trio("dla_up.ida_1.1", 128, h4, 2); trio("dla_up.ida_1.2", 128, h4, 2)
        trio("dla_up.ida_2.1", 64, h3, 2); trio("dla_up.ida_2.2", 64, h3, 2); trio("dla_up.ida_2.3", 64, h3, 2)
        trio("ida_up.1", 64, h3, 2); trio("ida_up.2", 64, h4, 4)
        b["hid"] = z(B, h2, h2, 256 * len(self.head_names))
        self.out = {h: torch.zeros(B, self.specs["head2." + h].n_valid, h2, h2, device=dev, dtype=torch.float32)
                    for h in self.head_names}
        self.buf = b
        self.inp = {"x": torch.zeros(B, 3, S, S, device=dev), "pre_img": torch.zeros(B, 3, S, S, device=dev),
                    "pre_hm": torch.zeros(B, 1, S, S, device=dev), "repro_hm": torch.zeros(B, 1, S, S, device=dev),
                    "pre_hm_cls": torch.zeros(B, 7, h2, h2, device=dev),
                    "repro_hm_cls": torch.zeros(B, 7, h2, h2, device=dev)}

    # ------------------------------------------------------------------ plan pieces
    def _basic_block(self, blk, x, ldx, xo, B, H, W, mid, res, ldres, reso, y, ldy, yo):
        c1, c2 = blk["c1"], blk["c2"]
        Ho, Wo = c1.out_hw(H, W)
        fo.conv_nhwc(c1, x, B, H, W, ldx, mid, c1.Cout, x_coff=xo)
        fo.conv_nhwc(c2, mid, B, Ho, Wo, c1.Cout, y, ldy, y_coff=yo, res=res, ldres=ldres, res_coff=reso)

    def _tree1(self, t, x, ldx, xo, B, H, W, stride, cin, cout, bot, ldbot, boto, resbuf, mid, cat, ldcat,
               out, ldout, outo):
        """Tree(levels=1): children already sit in `cat` beyond [0, 2*cout).  `bot` is the
        (pooled) input used for the projection / identity residual."""
        Ho, Wo = H // stride, W // stride
        if "project" in t:
            fo.conv_nhwc(t["project"], bot, B, Ho, Wo, ldbot, resbuf, cout, x_coff=boto)
            res, ldres, reso = resbuf, cout, 0
        else:
            res, ldres, reso = bot, ldbot, boto
        # x1 -> cat[:, cout:2cout], x2 -> cat[:, 0:cout]
        self._basic_block(t["tree1"], x, ldx, xo, B, H, W, mid, res, ldres, reso, cat, ldcat, cout)
        self._basic_block(t["tree2"], cat, ldcat, cout, B, Ho, Wo, mid, cat, ldcat, cout, cat, ldcat, 0)
        fo.conv_nhwc(t["root"], cat, B, Ho, Wo, ldcat, out, ldout, y_coff=outo)

    def _deform_conv(self, d, x, B, H, W, y):
        om = self.buf["om"]
        fo.conv_nhwc(d["om"], x, B, H, W, d["Cin"], om, 32)
        _lib.call("sgta_dcn_forward_nhwc", _lib.ptr(x), _lib.ptr(om), _lib.ptr(d["w"]), _lib.ptr(d["scale"]),
                  _lib.ptr(d["shift"]), _lib.ptr(y), B, d["Cin"], d["Cout"], H, W, self.mma, 1,
                  fo._dt(y.dtype), _lib.stream())

    def _ida_step(self, name, k, src, hs, skip, f):
        """layers[i] = node(up(proj(layers[i])) + layers[i-1])   (dla.py:571-577)"""
        s, b, B = self.specs, self.buf, self.B
        key = "%s.%d" % (name, k)
        proj, node = s["%s.proj_%d" % (name, k)], s["%s.node_%d" % (name, k)]
        self._deform_conv(proj, src, B, hs, hs, b[key + ".p"])
        c = proj["Cout"]
        fo.upsample_add(b[key + ".p"], s["%s.up_%d" % (name, k)], skip, c, b[key + ".s"], c, B, hs, hs, c, f)
        self._deform_conv(node, b[key + ".s"], B, hs * f, hs * f, b[key + ".n"])
        return b[key + ".n"]

    def _fuse_level(self, i, feats, pre_flat, rep_flat, Whm):
        B, sd = self.B, self.sd
        _, H, W, C = feats.shape
        pre_ids = fusion.window_ids(pre_flat[i], Whm, SCALE_LIST[i], self.kernel_list[i], H, W)
        cur_ids = fusion.window_ids(rep_flat[i], Whm, SCALE_LIST[i], self.kernel_list[i], H, W)
        pre_key = fo.gather_tokens_nhwc(feats[:B], C, pre_ids, C, H * W)
        cur_q = fo.gather_tokens_nhwc(feats[B:], C, cur_ids, C, H * W)
        out = pre_key
        if i <= 2:
            p = "transformer.%d.layers.0" % i
            a = p + ".cross_attn"
            heads = 8
            Kp = F.linear(pre_key, sd[a + ".w_k.weight"])            # K, V are shared by the 3 layers
            Vp = F.linear(pre_key, sd[a + ".w_v.weight"])
            scale = (Kp.shape[-1] // heads) ** 0.5
            pos = sd[a + ".pos_embed"] if self.use_pos else None
            q = cur_q
            for _ in range(3):                                        # one shared layer (dla.py:788-789)
                att = fusion.attention_core(F.linear(q, sd[a + ".w_q.weight"]), Kp, Vp, pos, heads, scale)
                q = F.layer_norm(F.linear(att, sd[a + ".fc.weight"], sd[a + ".fc.bias"]) + q, (C,),
                                 sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
                ffn = F.linear(F.relu(F.linear(q, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                               sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
                q = F.layer_norm(q + ffn, (C,), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])
            out = q
        c = "cat_layer.%d" % i
        rows = F.linear(F.relu(F.linear(torch.cat([out, cur_q], -1), sd[c + ".0.weight"], sd[c + ".0.bias"])),
                        sd[c + ".2.weight"], sd[c + ".2.bias"])
        fused = self.buf["fused%d" % i]
        fused.copy_(feats[B:])
        fo.scatter_tokens_nhwc(fused, C, cur_ids, rows, C, H * W)
        return fused

    # ------------------------------------------------------------------ the plan
    def _run(self):
        s, b, B, S = self.specs, self.buf, self.B, self.S
        B2 = 2 * B
        i = self.inp
        in4 = b["in4"]
        fo.nchw_to_nhwc(i["pre_img"], in4[:B], 4, 0)
        fo.nchw_to_nhwc(i["pre_hm"], in4[:B], 4, 3)
        fo.nchw_to_nhwc(i["x"], in4[B:], 4, 0)
        fo.nchw_to_nhwc(i["repro_hm"], in4[B:], 4, 3)
        fo.conv_nhwc(s["stem"], in4, B2, S, S, 4, b["f0"], 16, epi=fo.EPI_STEM)
        fo.conv_nhwc(s["level0"], b["f0"], B2, S, S, 16, b["l0"], 16)
        fo.conv_nhwc(s["level1"], b["l0"], B2, S, S, 16, b["l1"], 32)
        h1, h2, h3, h4, h5 = S // 2, S // 4, S // 8, S // 16, S // 32
        # level 2: Tree(1, 32->64, stride 2)
        fo.maxpool2(b["l1"], B2, h1, h1, 32, 32, b["bot2"], 32)
        self._tree1(s["level2"], b["l1"], 32, 0, B2, h1, h1, 2, 32, 64, b["bot2"], 32, 0, b["res2"], b["mid2"],
                    b["cat2"], 128, b["l2"], 64, 0)
        # levels 3, 4: Tree(2, c/2 -> c, stride 2, level_root): cat_b = [x2 | x1 | bottom | tree1 out]
        prev, hp, cp = b["l2"], h2, 64
        for lv, c, h in ((3, 128, h3), (4, 256, h4)):
            t = s["level%d" % lv]
            catb, ldb = b["cat%db" % lv], 3 * c + c // 2
            fo.maxpool2(prev, B2, hp, hp, cp, cp, catb, ldb, y_coff=2 * c)                # bottom -> children
            self._tree1(t["tree1"], prev, cp, 0, B2, hp, hp, 2, cp, c, catb, ldb, 2 * c, b["res%da" % lv],
                        b["mid%d" % lv], b["cat%da" % lv], 2 * c, catb, ldb, 2 * c + cp)     # x1 -> children
            self._tree1(t["tree2"], catb, ldb, 2 * c + cp, B2, h, h, 1, c, c, catb, ldb, 2 * c + cp, None,
                        b["mid%d" % lv], catb, ldb, b["l%d" % lv], c, 0)
            prev, hp, cp = b["l%d" % lv], h, c
        # level 5: Tree(1, 256->512, stride 2, level_root): cat = [x2 | x1 | bottom]
        fo.maxpool2(b["l4"], B2, h4, h4, 256, 256, b["cat5"], 1280, y_coff=1024)
        self._tree1(s["level5"], b["l4"], 256, 0, B2, h4, h4, 2, 256, 512, b["cat5"], 1280, 1024, b["res5"],
                    b["mid5"], b["cat5"], 1280, b["l5"], 512, 0)
        # structure-prior fusion
        pre_flat, rep_flat = {}, {}
        for K in set(self.K_list):
            p, r = fusion.topk_flat_index(i["pre_hm_cls"], K), fusion.topk_flat_index(i["repro_hm_cls"], K)
            for lv in range(6):
                if self.K_list[lv] == K:
                    pre_flat[lv], rep_flat[lv] = p, r
        fused = {}
        for lv in range(6):
            if lv < 2 and self.skip_dead_levels:
                continue                     # levels 0/1 never reach the output (DLAUp starts at level 2)
            fused[lv] = self._fuse_level(lv, b["l%d" % lv], pre_flat, rep_flat, h2)
        # DLAUp (dla.py:600-606)
        a5 = self._ida_step("dla_up.ida_0", 1, fused[5], h5, fused[4], 2)                  # 256 @ h4
        b4 = self._ida_step("dla_up.ida_1", 1, fused[4], h4, fused[3], 2)                  # 128 @ h3
        b5 = self._ida_step("dla_up.ida_1", 2, a5, h4, b4, 2)
        c3 = self._ida_step("dla_up.ida_2", 1, fused[3], h3, fused[2], 2)                  # 64 @ h2
        c4 = self._ida_step("dla_up.ida_2", 2, b4, h3, c3, 2)
        c5 = self._ida_step("dla_up.ida_2", 3, b5, h3, c4, 2)
        # IDAUp over [c5 (64@h2), b5 (128@h3), a5 (256@h4)]  (dla.py:1548-1552)
        y1 = self._ida_step("ida_up", 1, b5, h3, c5, 2)
        y2 = self._ida_step("ida_up", 2, a5, h4, y1, 4)
        # heads
        nh = len(self.head_names)
        fo.conv_nhwc(s["head0"], y2, B, h2, h2, 64, b["hid"], 256 * nh)
        for j, h in enumerate(self.head_names):
            fo.conv_nhwc(s["head2." + h], b["hid"], B, h2, h2, 256 * nh, self.out[h], 0, x_coff=256 * j,
                         epi=fo.EPI_NCHW)
        self.feat = y2

    # ------------------------------------------------------------------ public API
    def forward(self, x, pre_img, pre_hm, repro_hm, pre_hm_cls, repro_hm_cls):
        """Same arguments as the reference forward (NCHW fp32, device or pinned host tensors);
        returns [ {hm, reg, tracking} ] (static buffers, overwritten by the next call)."""
        for k, t in (("x", x), ("pre_img", pre_img), ("pre_hm", pre_hm), ("repro_hm", repro_hm),
                     ("pre_hm_cls", pre_hm_cls), ("repro_hm_cls", repro_hm_cls)):
            self.inp[k].copy_(t, non_blocking=True)
        with torch.no_grad():
            if not self._use_graph:
                self._run()
            else:
                if self.graph is None:
                    self._run()                                   # warm-up: attributes, allocator
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._run()
                    self.graph = g
                self.graph.replay()
        return [dict(self.out)]

    __call__ = forward

    def infer(self, *inputs):
        """forward -> sigmoid -> live decode, like SGTADetector.process (sgta_detector.py:881-927)."""
        out = dict(self.forward(*inputs)[0])
        if not self.fuse_sigmoid:
            out["hm"] = torch.sigmoid(out["hm"])
        return decode.dream_generic_decode(out, K=out["hm"].shape[1], opt=self.opt)
