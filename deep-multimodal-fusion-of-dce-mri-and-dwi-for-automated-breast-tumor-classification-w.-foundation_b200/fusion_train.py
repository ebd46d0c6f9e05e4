"""Fusion-head fine-tuning step on the B200 kernels (BASELINE config C5, frozen-encoder phase).

What the reference does in that phase (code/train_fusion.py:203-321 with the optimiser factory of
code/selector_helpers.py:356-520, `backbone_freeze_on_start`): both encoders are frozen, the fusion head is the one
trainable parameter group, and every step is  encoders -> FusionModel.forward -> loss -> backward -> AdamW.
This module is the classification-objective part of that step:

    loss = Soft(Weighted)FocalLoss(logits, LabelSmoothing(logits, labels))       train_fusion.py:238-242

The logits of FusionModel.forward (code/model_module.py:919-1000) depend on the encoder maps only through their
4x4-pooled tokens (proj_in_* are bias-free 1x1 convolutions, GAP and the bilinear up-sample are linear), so one
pooling pass over each f3 map is the only map-sized work; forward, backward and the weight gradients then run as
fp32 kernels on [B*16, C] token matrices (csrc/train_ops.cu).  The parameters that receive a gradient - the same
20 tensors torch autograd finds on the reference module (tests/golden/train_head.npz) - live in ONE flat fp32
buffer (the nn.Parameters are views of it, so state_dict / load_state_dict keep working), their gradients in a
second one: a data-parallel step is one all-reduce of that buffer (NCCL over NVLink on the GPUs, gloo in the CPU
tests of the host logic) and one fused AdamW launch.  Parameters off the logits path (mask head, reconstruction
head, projector, the reference's dead reduce/refine branch) get no gradient and are left untouched, exactly like
torch.optim.AdamW skips parameters whose .grad is None.

Optional second term (lambda_mask > 0): the reference's mask dice term, lambda_mask * (dice(dwi mask) + dice(dce
mask) + dice(fused mask)) / 3 (train_fusion.py:245-255, loss.py:45-62).  Only the fused mask reaches the head's
parameters; at the mask size MaskHeadResize is two 1x1 convolutions with nothing in between, so the fused mask logits
are linear in fused_refined and cost two more passes over each f3 map (a per-case 512-vector dot product forward, a
dmask-weighted pixel sum backward) - no 128-channel full-resolution map is ever materialised.  mask_head.pre / .out
then join the trainable set (24 tensors).

Not built: the reconstruction / mimic terms of the reference's total loss (they need the training-mode BatchNorm
backward of the reconstruction and projector heads) and unfrozen encoders.  There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

import b200_native as nat
from model_module import _as_nhwc_bf16, _bilinear_axis_weights

__all__ = ["FusionHeadTrainer", "flat_views", "flat_size", "average_gradients"]


ALIGN = 64  # elements: every tensor starts on a 256-byte boundary of the flat buffers (kernels that consume the
            # parameters elsewhere - the inference path - use 16-byte vector loads)


def flat_size(tensors, align=ALIGN):
    """Elements of a flat buffer holding `tensors`, each padded to a multiple of `align` elements."""
    return sum((t.numel() + align - 1) // align * align for t in tensors)


def flat_views(tensors, flat, align=ALIGN):
    """Views of `flat` with the shapes of `tensors`, in order, each starting on an `align`-element boundary (the
    layout of the parameter, gradient and moment buffers; the padding stays zero)."""
    out, off = [], 0
    for t in tensors:
        n = t.numel()
        out.append(flat[off:off + n].view(t.shape))
        off += (n + align - 1) // align * align
    return out


def average_gradients(flat_grads, group=None):
    """Data-parallel exchange of one step: SUM all-reduce of the flat gradient buffer; returns the factor the
    optimiser kernel must apply (1 / world size).  One collective per step."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def _split_k(m, n, k):
    tiles = ((m + 63) // 64) * ((n + 63) // 64)
    return max(1, min((2 * 148 + tiles - 1) // tiles, max(1, k // 64)))


class FusionHeadTrainer(torch.optim.Optimizer):
    """Owns the flat parameter / gradient / AdamW-moment buffers of a `model_module.FusionModel` and runs its
    fine-tuning step.  `lr`, `betas`, `eps`, `weight_decay` as torch.optim.AdamW (code/selector_helpers.py:222-229);
    `smoothing` = label_smoothing_alpha, `gamma` / `class_weights` as Soft(Weighted)FocalLoss
    (code/selector_helpers.py:14-46).

    It IS a `torch.optim.Optimizer` (one parameter group, the fusion head): torch's learning-rate schedulers - the
    ones the reference's `_build_scheduler` creates (code/selector_helpers.py:692-728) - and harnesses that expect an
    optimizer from `configure_optimizers` drive it unchanged; `step()` launches the fused AdamW kernel with the
    group's current `lr` / `weight_decay`, `zero_grad()` clears the flat gradient buffer (the `.grad` views are never
    dropped, whatever `set_to_none` says)."""

    def __init__(self, fusion_model, *, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=4e-5, smoothing=0.1,
                 gamma=1.5, class_weights=None, lambda_mask=0.0, mask_loss_type="dice", process_group=None):
        fm = fusion_model
        if isinstance(fm.proj_in_dwi, nn.Identity) or isinstance(fm.proj_in_dce, nn.Identity):
            raise NotImplementedError("encoder channels == fusion_channels (identity proj_in) is not built")
        hp, wp = fm.token_pool
        if fm.use_cross_attention and (hp * wp > 32 or fm.fusion_channels % fm.mha_heads != 0 or
                                       fm.fusion_channels // fm.mha_heads > 128):
            raise NotImplementedError("token_pool / head size outside the attention kernel's range")
        self.model = fm
        self.smoothing, self.gamma = float(smoothing), float(gamma)
        self.class_weights = class_weights
        self.lambda_mask = float(lambda_mask)  # > 0: + lambda_mask * mean of the three mask loss terms
        if mask_loss_type not in ("dice", "dice_bce"):  # mask_criterion_selector, code/selector_helpers.py:95-114
            raise ValueError(f"Invalid mask loss: {mask_loss_type}")
        self.mask_loss_type = mask_loss_type
        self.group = process_group
        self.step_count = 0
        named = [("proj_in_dwi.weight", fm.proj_in_dwi.weight), ("proj_in_dce.weight", fm.proj_in_dce.weight),
                 ("gating.fc.weight", fm.gating.fc.weight), ("gating.fc.bias", fm.gating.fc.bias)]
        if fm.use_cross_attention:
            ca = fm.cross_attn_block
            named += [("cross_attn_block.cross_attn.in_proj_weight", ca.cross_attn.in_proj_weight),
                      ("cross_attn_block.cross_attn.in_proj_bias", ca.cross_attn.in_proj_bias),
                      ("cross_attn_block.cross_attn.out_proj.weight", ca.cross_attn.out_proj.weight),
                      ("cross_attn_block.cross_attn.out_proj.bias", ca.cross_attn.out_proj.bias),
                      ("cross_attn_block.attn_ffn.0.weight", ca.attn_ffn[0].weight),
                      ("cross_attn_block.attn_ffn.0.bias", ca.attn_ffn[0].bias),
                      ("cross_attn_block.attn_ffn.1.weight", ca.attn_ffn[1].weight),
                      ("cross_attn_block.attn_ffn.1.bias", ca.attn_ffn[1].bias),
                      ("cross_attn_block.attn_ffn.3.weight", ca.attn_ffn[3].weight),
                      ("cross_attn_block.attn_ffn.3.bias", ca.attn_ffn[3].bias)]
        if fm.fusion_se is not None:
            se = fm.fusion_se
            named += [("fusion_se.fc.1.weight", se.fc[1].weight), ("fusion_se.fc.1.bias", se.fc[1].bias),
                      ("fusion_se.fc.3.weight", se.fc[3].weight), ("fusion_se.fc.3.bias", se.fc[3].bias)]
        named += [("classifier.2.weight", fm.classifier[2].weight), ("classifier.2.bias", fm.classifier[2].bias)]
        if self.lambda_mask > 0:  # the fused mask head joins the trainable set (its logits enter the loss)
            mh = fm.mask_head
            named += [("mask_head.pre.weight", mh.pre.weight), ("mask_head.pre.bias", mh.pre.bias),
                      ("mask_head.out.weight", mh.out.weight), ("mask_head.out.bias", mh.out.bias)]
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.numel = sum(p.numel() for p in self.params)   # trainable parameters
        self.flat_numel = flat_size(self.params)            # elements of the flat buffers (with alignment padding)
        self._flat = None
        self._ws = {}
        super().__init__(self.params, dict(lr=float(lr), betas=tuple(betas), eps=float(eps),
                                           weight_decay=float(weight_decay)))

    # hyper-parameters live in the (single) parameter group, where torch's schedulers read and write them
    def _hp(name):  # noqa: N805
        return property(lambda self: self.param_groups[0][name],
                        lambda self, value: self.param_groups[0].__setitem__(name, value))

    lr, betas, eps, weight_decay = _hp("lr"), _hp("betas"), _hp("eps"), _hp("weight_decay")
    del _hp

    # ------------------------------------------------------------------------------------------ buffers ----
    def _bind(self):
        """(Re)create the flat buffers on the parameters' device and make the parameters views of them."""
        dev = self.params[0].device
        if dev.type != "cuda":
            raise nat.B200NativeError("FusionHeadTrainer needs the fusion model on a CUDA device (no CPU path)")
        bound = self._flat is not None and self._flat["p"].device == dev
        if bound:
            for p, view in zip(self.params, flat_views(self.params, self._flat["p"])):
                bound = bound and p.data_ptr() == view.data_ptr() and p.dtype == torch.float32
        if bound:
            return self._flat
        old = self._flat
        n = self.flat_numel
        flat = {"p": torch.zeros(n, dtype=torch.float32, device=dev),
                "g": torch.zeros(n + 1, dtype=torch.float32, device=dev),  # last element: the loss
                "m": torch.zeros(n, dtype=torch.float32, device=dev),
                "v": torch.zeros(n, dtype=torch.float32, device=dev)}
        if old is not None:  # the model was moved / reloaded: keep the optimiser state
            flat["m"].copy_(old["m"])
            flat["v"].copy_(old["v"])
        for p, view in zip(self.params, flat_views(self.params, flat["p"])):
            view.copy_(p.data.float())
            p.data = view
        self.grads = flat_views(self.params, flat["g"])
        for p, g in zip(self.params, self.grads):
            p.grad = g
        cw = self.class_weights
        flat["cw"] = None if cw is None else torch.as_tensor(cw, dtype=torch.float32).to(dev).contiguous()
        self._flat = flat
        self._ws = {}
        return flat

    def _workspace(self, B, H, W):
        key = (B, H, W)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        fm = self.model
        dev = self.params[0].device
        hp, wp = fm.token_pool
        T, C = hp * wp, fm.fusion_channels
        R = B * T
        Cm = fm.fusion_se.fc[1].weight.shape[0] if fm.fusion_se is not None else 0
        in_dim = 2 * C + (2 if fm.use_mask_attention else 0)
        K = fm.num_classes

        def z(*shape):
            return torch.empty(shape, dtype=torch.float32, device=dev)

        ws = {"Xd": z(R, fm.dwi_ch), "Xc": z(R, fm.dce_ch), "Td": z(R, C), "Tc": z(R, C), "Q": z(R, C),
              "KV": z(R, 2 * C), "P": z(B, fm.mha_heads, T, T), "CTX": z(R, C), "AO": z(R, C), "LN": z(R, C),
              "mean": z(R), "rstd": z(R), "H1": z(R, C), "G1": z(R, C), "LOW": z(R, C),
              "logits": z(B, K), "gating": z(B, 2), "dlogits": z(B, K), "zvec": z(B, C), "gf": z(B, C),
              "h": z(B, max(Cm, 1)), "da1": z(B, max(Cm, 1)), "da2": z(B, C), "gx": z(B, in_dim), "dgl": z(B, 2),
              "dpd": z(B, C), "dpc": z(B, C), "dLOW": z(R, C), "dG1": z(R, C), "dH1": z(R, C), "dLN": z(R, C),
              "dAO": z(R, C), "tmp": z(R, C), "dCTX": z(R, C), "dQ": z(R, C), "dKV": z(R, 2 * C), "dTd": z(R, C),
              "dTc": z(R, C)}
        if self.lambda_mask > 0:
            ws.update({"v": z(1, C), "gate": z(B, C), "u": z(B, C), "omd": z(B, fm.dwi_ch), "omc": z(B, fm.dce_ch),
                       "Dd": z(B, H * W), "Dc": z(B, H * W), "m": z(B, 1, fm.mask_size, fm.mask_size), "dm": z(B, H * W),
                       "q": z(B, T),
                       "sd": z(B, fm.dwi_ch), "sc": z(B, fm.dce_ch), "tmpd": z(B, C), "tmpc": z(B, C), "dug": z(B, C),
                       "aud": z(B, C), "auc": z(B, C), "dv": z(C + 1)})  # dv[C] = dc0
        ah, aw = _bilinear_axis_weights(hp, H), _bilinear_axis_weights(wp, W)
        ws["up"] = torch.tensor([ah[i] * aw[j] for i in range(hp) for j in range(wp)], dtype=torch.float32,
                                device=dev)
        self._ws = {key: ws}  # one batch geometry at a time
        return ws

    # ---------------------------------------------------------------------------------------- the step ----
    def zero_grad(self, set_to_none=False):
        self._bind()["g"].zero_()

    def loss_and_grads(self, f3_dwi, f3_dce, dwi_mask_pred, dce_mask_pred, labels, masks=None):
        """Forward + backward of the objective on one batch; ACCUMULATES into the flat gradient buffer (call
        zero_grad first).  f3_* are the encoders' deepest maps ([B,C,H,W]-shaped, bf16 channels-last as the B200
        encoders emit them), *_mask_pred their mask logits, `masks` the [B,1,H,W] target masks (needed when
        lambda_mask > 0).  Returns (loss, logits) - device tensors, the loss being this rank's batch mean; with the
        mask term `self.fused_mask_logits` holds the fused mask logits of the step."""
        fm = self.model
        flat = self._bind()
        f3d, f3c = _as_nhwc_bf16(f3_dwi), _as_nhwc_bf16(f3_dce)
        B, H, W, _ = f3d.shape
        if B == 0:
            raise ValueError("empty batch: the mean-reduced loss is undefined")
        if f3c.shape[:3] != f3d.shape[:3]:
            raise ValueError("the two encoders' maps must have the same batch and spatial size")
        hp, wp = fm.token_pool
        # equal bins: GAP(p) is the mean of the tokens.  Otherwise (14 x 14 ViT maps -> 4 x 4 overlapping bins) the
        # pooled vectors get their own pass over the maps (per-case channel sums) and their own projection.
        cross = bool(fm.use_cross_attention)
        gap_rows = bool(H % hp or W % wp) or not cross  # (without cross-attention the tokens are not needed at all)
        if fm.use_mask_attention and (dwi_mask_pred is None or dce_mask_pred is None):
            raise RuntimeError("use_mask_attention needs both encoder mask predictions")
        T, C, NH = hp * wp, fm.fusion_channels, fm.mha_heads
        R = B * T
        use_mask_term = self.lambda_mask > 0
        if use_mask_term:
            if masks is None:
                raise ValueError("lambda_mask > 0 needs the target masks")
            ms = fm.mask_size
            if H != W or H in (64, 128, 256, 512):
                raise NotImplementedError("the mask term is built for MaskHeadResize's identity (32) and interpolation "
                                          "dispatches; 64 / 128 / 256 / 512-pixel maps go through GELU convolutions")
            if tuple(masks.shape[-2:]) != (ms, ms) or masks.shape[0] != B:
                raise ValueError("target masks must be [B,1,H,W] at the mask size")
        ws = self._workspace(B, H, W)
        par = dict(zip(self.names, self.params))
        grd = dict(zip(self.names, self.grads))
        ca = "cross_attn_block."
        Wd, Wc = par["proj_in_dwi.weight"].view(C, -1), par["proj_in_dce.weight"].view(C, -1)
        if cross:
            Win, b_in = par[ca + "cross_attn.in_proj_weight"], par[ca + "cross_attn.in_proj_bias"]
            Wo, bo = par[ca + "cross_attn.out_proj.weight"], par[ca + "cross_attn.out_proj.bias"]
            lnw, lnb = par[ca + "attn_ffn.0.weight"], par[ca + "attn_ffn.0.bias"]
            W1, b1 = par[ca + "attn_ffn.1.weight"], par[ca + "attn_ffn.1.bias"]
            W2, b2 = par[ca + "attn_ffn.3.weight"], par[ca + "attn_ffn.3.bias"]

        # ---- forward on pooled tokens ----
        if gap_rows:
            for k in ("Gd", "Gc", "Pd", "Pc"):
                if k not in ws:
                    ws[k] = torch.empty((B, fm.dwi_ch if k == "Gd" else fm.dce_ch if k == "Gc" else C),
                                        dtype=torch.float32, device=f3d.device)
            nat.channel_sums(f3d, ws["Gd"])
            nat.channel_sums(f3c, ws["Gc"])
            nat.sgemm(ws["Gd"], Wd, ws["Pd"], trans_b=True)
            nat.sgemm(ws["Gc"], Wc, ws["Pc"], trans_b=True)
        if cross:
            nat.fusion_tokens(f3d, hp, wp, ws["Xd"])
            nat.fusion_tokens(f3c, hp, wp, ws["Xc"])
            nat.sgemm(ws["Xd"], Wd, ws["Td"], trans_b=True)
            nat.sgemm(ws["Xc"], Wc, ws["Tc"], trans_b=True)
            nat.sgemm(ws["Td"], Win[:C], ws["Q"], trans_b=True, bias=b_in[:C])
            nat.sgemm(ws["Tc"], Win[C:], ws["KV"], trans_b=True, bias=b_in[C:])
            Kt, Vt = ws["KV"][:, :C], ws["KV"][:, C:]
            nat.mha_fwd(ws["Q"], Kt, Vt, B, NH, ws["P"], ws["CTX"])
            nat.sgemm(ws["CTX"], Wo, ws["AO"], trans_b=True, bias=bo)
            nat.ln_fwd(ws["AO"], lnw, lnb, fm.cross_attn_block.attn_ffn[0].eps, ws["LN"], ws["mean"], ws["rstd"])
            nat.sgemm(ws["LN"], W1, ws["G1"], trans_b=True, bias=b1, pre=ws["H1"], act=1)
            nat.sgemm(ws["G1"], W2, ws["LOW"], trans_b=True, bias=b2, res=ws["AO"])

        # ---- per-case tail, loss, and the backward of the tail ----
        a = nat.HeadTrain()
        a.C, a.T, a.num_classes = C, T, fm.num_classes
        a.use_mask_attention, a.use_se = int(fm.use_mask_attention), int(fm.fusion_se is not None)
        a.smoothing, a.gamma, a.loss_scale = self.smoothing, self.gamma, 1.0 / B
        keep = []

        def ptr(t):
            keep.append(t)
            return t.data_ptr()

        if flat["cw"] is not None:
            a.class_weights = ptr(flat["cw"])
        a.tok_dwi, a.tok_dce = ptr(ws["Td"]), ptr(ws["Tc"])  # (not read when the pooled vectors are given)
        if cross:
            a.lowres = ptr(ws["LOW"])
        if gap_rows:
            a.pvec_dwi, a.pvec_dce, a.pvec_scale = ptr(ws["Pd"]), ptr(ws["Pc"]), 1.0 / (H * W)
        md = dwi_mask_pred.contiguous().float() if dwi_mask_pred is not None else None
        mc = dce_mask_pred.contiguous().float() if dce_mask_pred is not None else None
        if fm.use_mask_attention:
            a.npix_mask = md[0].numel()
            a.mask_dwi, a.mask_dce = ptr(md), ptr(mc)
        lab = labels.to(device=f3d.device, dtype=torch.int64).contiguous()
        if lab.numel() != B:
            raise ValueError("one label per case is required")
        a.labels = ptr(lab)
        a.gate_w, a.gate_b = ptr(par["gating.fc.weight"]), ptr(par["gating.fc.bias"])
        a.up_coef = ptr(ws["up"])
        if fm.fusion_se is not None:
            a.se_mid = par["fusion_se.fc.1.weight"].shape[0]
            a.se_w1, a.se_b1 = ptr(par["fusion_se.fc.1.weight"]), ptr(par["fusion_se.fc.1.bias"])
            a.se_w2, a.se_b2 = ptr(par["fusion_se.fc.3.weight"]), ptr(par["fusion_se.fc.3.bias"])
            a.h_out, a.da1_out, a.da2_out = ptr(ws["h"]), ptr(ws["da1"]), ptr(ws["da2"])
        a.cls_w, a.cls_b = ptr(par["classifier.2.weight"]), ptr(par["classifier.2.bias"])
        loss = flat["g"][self.flat_numel:]
        a.loss_out, a.logits_out, a.gating_out = ptr(loss), ptr(ws["logits"]), ptr(ws["gating"])
        a.dlogits_out, a.z_out, a.gf_out = ptr(ws["dlogits"]), ptr(ws["zvec"]), ptr(ws["gf"])
        a.gx_out, a.dgl_out = ptr(ws["gx"]), ptr(ws["dgl"])
        a.dpd_out, a.dpc_out = ptr(ws["dpd"]), ptr(ws["dpc"])
        if cross:
            a.dlowres_out = ptr(ws["dLOW"])
        if use_mask_term:
            # fused mask logit = c0 + v . fused_refined (pre and out are both 1x1, nothing in between at this size)
            pre_w = par["mask_head.pre.weight"].view(par["mask_head.pre.weight"].shape[0], C)
            pre_b, out_w = par["mask_head.pre.bias"], par["mask_head.out.weight"].view(1, -1)
            out_b = par["mask_head.out.bias"]
            nat.sgemm(out_w, pre_w, ws["v"])
            a.mask_v, a.gate_out, a.u_out = ptr(ws["v"]), ptr(ws["gate"]), ptr(ws["u"])
            a.forward_only = 1
            nat.head_loss(a, B)                                   # gating, SE gate, u = v * gate
            a.forward_only = 0
            nat.sgemm(ws["u"], Wd, ws["omd"])                     # omega = W^T u: one 512-vector per case
            nat.sgemm(ws["u"], Wc, ws["omc"])
            nat.mask_dot(f3d, ws["omd"], ws["Dd"])                # second pass over the maps
            nat.mask_dot(f3c, ws["omc"], ws["Dc"])
            tgt = masks.to(device=f3d.device, dtype=torch.float32).contiguous()
            ws["dv"].zero_()
            if md is None or mc is None or md[0].numel() != ms * ms or mc[0].numel() != ms * ms:
                raise ValueError("the mask term needs both encoder mask predictions at the mask size "
                                 "(train_fusion.py:249-251)")
            nat.mask_dice(ws["Dd"], ws["Dc"], ws["gating"], ws["u"], ws["LOW"] if cross else None, pre_b, out_w, out_b, tgt, md, mc,
                          H, W, ms, ms, hp, wp, self.lambda_mask / (3.0 * B), 1e-6, int(self.mask_loss_type == "dice_bce"),
                          ws["m"], ws["dm"], ws["q"], ws["dv"][C:], loss)
            nat.mask_wsum(f3d, ws["dm"], ws["sd"])                # third pass: dmask-weighted pixel sums
            nat.mask_wsum(f3c, ws["dm"], ws["sc"])
            nat.sgemm(ws["sd"], Wd, ws["tmpd"], trans_b=True)
            nat.sgemm(ws["sc"], Wc, ws["tmpc"], trans_b=True)
            a.mk_tmpd, a.mk_tmpc, a.mk_q = ptr(ws["tmpd"]), ptr(ws["tmpc"]), ptr(ws["q"])
            a.dug_out, a.aud_out, a.auc_out = ptr(ws["dug"]), ptr(ws["aud"]), ptr(ws["auc"])
            self.fused_mask_logits = ws["m"]
        nat.head_loss(a, B)
        if use_mask_term:
            nat.colsum(ws["dug"], ws["dv"][:C])
            nat.mask_head_grads(ws["dv"][:C], ws["dv"][C:], pre_w, pre_b, out_w, grd["mask_head.pre.weight"],
                                grd["mask_head.pre.bias"], grd["mask_head.out.weight"], grd["mask_head.out.bias"])

        def wgrad(dy, x, name, rows=None):
            g = grd[name] if rows is None else grd[name][rows]
            g = g.view(g.shape[0], -1)
            nat.sgemm(dy, x, g, trans_a=True, beta=1, split_k=_split_k(g.shape[0], g.shape[1], dy.shape[0]))

        def bgrad(dy, name, rows=None):
            nat.colsum(dy, grd[name] if rows is None else grd[name][rows])

        wgrad(ws["dlogits"], ws["zvec"], "classifier.2.weight")
        bgrad(ws["dlogits"], "classifier.2.bias")
        if fm.fusion_se is not None:
            wgrad(ws["da1"], ws["gf"], "fusion_se.fc.1.weight")
            bgrad(ws["da1"], "fusion_se.fc.1.bias")
            wgrad(ws["da2"], ws["h"], "fusion_se.fc.3.weight")
            bgrad(ws["da2"], "fusion_se.fc.3.bias")
        wgrad(ws["dgl"], ws["gx"], "gating.fc.weight")
        bgrad(ws["dgl"], "gating.fc.bias")

        # ---- backward through the cross-attention block ----
        if cross:
            dLOW = ws["dLOW"]
            wgrad(dLOW, ws["G1"], ca + "attn_ffn.3.weight")
            bgrad(dLOW, ca + "attn_ffn.3.bias")
            nat.sgemm(dLOW, W2, ws["dG1"])
            nat.gelu_bwd(ws["H1"], ws["dG1"], ws["dH1"])
            wgrad(ws["dH1"], ws["LN"], ca + "attn_ffn.1.weight")
            bgrad(ws["dH1"], ca + "attn_ffn.1.bias")
            nat.sgemm(ws["dH1"], W1, ws["dLN"])
            nat.ln_bwd(ws["AO"], ws["dLN"], dLOW, lnw, ws["mean"], ws["rstd"], ws["dAO"], ws["tmp"])
            bgrad(ws["tmp"], ca + "attn_ffn.0.weight")
            bgrad(ws["dLN"], ca + "attn_ffn.0.bias")
            wgrad(ws["dAO"], ws["CTX"], ca + "cross_attn.out_proj.weight")
            bgrad(ws["dAO"], ca + "cross_attn.out_proj.bias")
            nat.sgemm(ws["dAO"], Wo, ws["dCTX"])
            dK, dV = ws["dKV"][:, :C], ws["dKV"][:, C:]
            nat.mha_bwd(ws["Q"], Kt, Vt, ws["P"], ws["dCTX"], B, NH, ws["dQ"], dK, dV)
            q_rows, kv_rows = slice(0, C), slice(C, 3 * C)
            wgrad(ws["dQ"], ws["Td"], ca + "cross_attn.in_proj_weight", q_rows)
            bgrad(ws["dQ"], ca + "cross_attn.in_proj_bias", q_rows)
            wgrad(ws["dKV"], ws["Tc"], ca + "cross_attn.in_proj_weight", kv_rows)
            bgrad(ws["dKV"], ca + "cross_attn.in_proj_bias", kv_rows)
            if gap_rows:
                nat.sgemm(ws["dQ"], Win[:C], ws["dTd"])
                nat.sgemm(ws["dKV"], Win[C:], ws["dTc"])
            else:
                nat.sgemm(ws["dQ"], Win[:C], ws["dTd"], res=ws["dpd"], res_div=T)
                nat.sgemm(ws["dKV"], Win[C:], ws["dTc"], res=ws["dpc"], res_div=T)
            wgrad(ws["dTd"], ws["Xd"], "proj_in_dwi.weight")
            wgrad(ws["dTc"], ws["Xc"], "proj_in_dce.weight")
        if gap_rows:  # the pooled vectors' gradients reach proj_in_* through the channel sums, not the tokens
            wgrad(ws["dpd"], ws["Gd"], "proj_in_dwi.weight")
            wgrad(ws["dpc"], ws["Gc"], "proj_in_dce.weight")
        if use_mask_term:  # the full-resolution path of the mask logits into proj_in_*
            wgrad(ws["aud"], ws["sd"], "proj_in_dwi.weight")
            wgrad(ws["auc"], ws["sc"], "proj_in_dce.weight")
        return loss, ws["logits"]

    def step(self, closure=None):
        """Gradient all-reduce (when torch.distributed is initialised) + one fused AdamW launch."""
        if closure is not None:
            raise NotImplementedError("closures are not supported: the step has no autograd graph to re-evaluate")
        flat = self._bind()
        scale = average_gradients(flat["g"], self.group)
        self.step_count += 1
        nat.adamw(flat["p"], flat["g"][:self.flat_numel], flat["m"], flat["v"], lr=float(self.lr), betas=self.betas,
                  eps=self.eps, weight_decay=self.weight_decay, step=self.step_count, grad_scale=scale)
        for p in self.params:  # the kernel wrote through raw pointers: tell torch (packed-weight caches key on it)
            torch.autograd.graph.increment_version(p)
        return flat["g"][self.flat_numel:] * scale  # the loss averaged over ranks

    def train_step(self, f3_dwi, f3_dce, dwi_mask_pred, dce_mask_pred, labels, masks=None):
        """zero_grad -> loss_and_grads -> all-reduce -> AdamW.  Returns (loss averaged over ranks, logits)."""
        self.zero_grad()
        _, logits = self.loss_and_grads(f3_dwi, f3_dce, dwi_mask_pred, dce_mask_pred, labels, masks)
        return self.step(), logits

    # torch.optim-like surface for harnesses that treat the object returned by configure_optimizers as one
    def state_dict(self):
        """torch.optim.AdamW-shaped: {"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [{..., "params":
        [0..n-1]}]} (what generic checkpoint code and `torch.optim.Optimizer.load_state_dict` consumers index), plus
        "names" / "flat_numel" so that a state saved for another parameter set is refused with a clear message."""
        flat = self._bind()
        step = torch.tensor(float(self.step_count))
        state = {i: {"step": step.clone(), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
                 for i, (m, v) in enumerate(zip(flat_views(self.params, flat["m"]), flat_views(self.params, flat["v"])))}
        group = {k: self.param_groups[0][k] for k in ("lr", "betas", "eps", "weight_decay")}
        group["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [group], "names": list(self.names), "flat_numel": self.flat_numel}

    def load_state_dict(self, sd):
        """Accepts the dict above or the flat form earlier checkpoints hold ({"step", "names", "exp_avg", "exp_avg_sq",
        "hyper"})."""
        flat = self._bind()
        if "names" in sd and list(sd["names"]) != list(self.names):
            raise ValueError("optimizer state was saved for a different parameter set")
        if "state" in sd:
            if len(sd["state"]) != len(self.params):
                raise ValueError(f"optimizer state holds {len(sd['state'])} tensors, this trainer {len(self.params)}")
            ms, vs = flat_views(self.params, flat["m"]), flat_views(self.params, flat["v"])
            for i, (m, v) in enumerate(zip(ms, vs)):
                st = sd["state"][i]
                if tuple(st["exp_avg"].shape) != tuple(m.shape):
                    raise ValueError(f"optimizer state {i} ({self.names[i]}): shape {tuple(st['exp_avg'].shape)} != {tuple(m.shape)}")
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
            self.step_count = int(float(sd["state"][0]["step"])) if len(sd["state"]) else 0
            hyper = {k: v for k, v in sd["param_groups"][0].items() if k != "params"}
        else:
            if sd["exp_avg"].numel() != self.flat_numel:
                raise ValueError(f"optimizer state holds {sd['exp_avg'].numel()} moment elements, this trainer {self.flat_numel}")
            self.step_count = int(sd["step"])
            flat["m"].copy_(sd["exp_avg"])
            flat["v"].copy_(sd["exp_avg_sq"])
            hyper = sd.get("hyper", {})
        self.param_groups[0].update(hyper)
