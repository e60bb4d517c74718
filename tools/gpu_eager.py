"""The hot path in EAGER PyTorch on the GPU (cuBLAS GEMMs + SDPA attention + stock elementwise kernels): the
"reference on the same B200" bar of SURVEY.md 8d / BASELINE.md 4 ("also timed").  bench.py times it next to the
product arm as ``gpu_eager_baseline``; none of libpio_sm100 is used here and nothing here is used by the product.

It is the reference's arithmetic with the reference's libraries, minus its host-side pathologies: boxes are pooled with
one batched matmul against a [B,R,P] weight tensor built on the device (the reference loops over boxes in Python with
>= 4 device->host syncs per box, src/bbox_utils.py:30-44, which would only measure the PCIe round trip), and the bank
is normalised once instead of per call.  ``use_cache=False`` keeps the reference's decode algorithm (the whole growing
sequence re-run at every step, decap.py:130-155); ``use_cache=True`` is the same model with a KV cache.
Device-agnostic: tests/test_bench_cpu.py pins it against the CPU oracle on small inputs.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

D, HEADS, DEPTH, PATCH, NREG = 768, 12, 12, 14, 4
N_LAYER, N_HEAD = 4, 4


class EagerPipeline:
    def __init__(self, vit_w: Dict[str, torch.Tensor], dec_w: Dict[str, torch.Tensor], bank: Optional[torch.Tensor], device,
                 autocast_bf16: bool = False):
        self.dev = torch.device(device)
        self.vw = {k: v.to(self.dev) for k, v in vit_w.items()}
        self.dw = {k: v.to(self.dev) for k, v in dec_w.items()}
        self.bank = None
        if bank is not None:
            b = bank[bank.norm(dim=-1) != 0].to(self.dev)            # im2txtprojection.py:345
            self.bank = b
            self.bank_n = b / b.norm(dim=-1, keepdim=True)           # (:367, hoisted out of the call)
            if autocast_bf16:
                self.bank, self.bank_n = self.bank.bfloat16(), self.bank_n.bfloat16()
        self.bf16 = autocast_bf16 and self.dev.type == "cuda"
        self._pe = {}

    def _ctx(self):
        return torch.autocast("cuda", dtype=torch.bfloat16) if self.bf16 else torch.autocast("cpu", enabled=False)

    # ---------------------------------------------------------------- DINOv2 ViT-B/14-reg (dinov2/models/vision_transformer.py)
    def _pos(self, g):
        if g not in self._pe:
            pe = self.vw["pos_embed"].float()
            n = pe.shape[1] - 1
            if g * g != n:
                m = int(math.sqrt(n))
                pp = pe[:, 1:].reshape(1, m, m, -1).permute(0, 3, 1, 2)
                pp = F.interpolate(pp, size=(g, g), mode="bicubic", antialias=True, align_corners=False)
                pe = torch.cat([pe[:, :1], pp.permute(0, 2, 3, 1).reshape(1, g * g, -1)], 1)
            self._pe[g] = pe
        return self._pe[g]

    @torch.no_grad()
    def vit(self, imgs):
        w = self.vw
        with self._ctx():
            B, _, H, _ = imgs.shape
            g = H // PATCH
            x = F.conv2d(imgs, w["patch_embed.proj.weight"], w["patch_embed.proj.bias"], stride=PATCH).flatten(2).transpose(1, 2)
            x = torch.cat([w["cls_token"].expand(B, -1, -1), x.float()], 1) + self._pos(g)
            x = torch.cat([x[:, :1], w["register_tokens"].expand(B, -1, -1), x[:, 1:]], 1)
            N = x.shape[1]
            qkv = None
            for i in range(DEPTH):
                p = f"blocks.{i}."
                h = F.layer_norm(x, (D,), w[p + "norm1.weight"], w[p + "norm1.bias"], eps=1e-6)
                qkv = F.linear(h, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"])
                t = qkv.reshape(B, N, 3, HEADS, D // HEADS).permute(2, 0, 3, 1, 4)
                o = F.scaled_dot_product_attention(t[0], t[1], t[2]).transpose(1, 2).reshape(B, N, D)
                x = x + w[p + "ls1.gamma"] * F.linear(o, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"]).float()
                h = F.layer_norm(x, (D,), w[p + "norm2.weight"], w[p + "norm2.bias"], eps=1e-6)
                h = F.gelu(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"]))
                x = x + w[p + "ls2.gamma"] * F.linear(h, w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"]).float()
            xn = F.layer_norm(x.float(), (D,), w["norm.weight"], w["norm.bias"], eps=1e-6)
            # CLS attention map of the last block: softmax_j(<q_cls, k_j>/128)  (dino_extraction.py:24-34)
            q_cls, k = qkv[:, 0, :D].float(), qkv[:, 1 + NREG:, D:2 * D].float()
            attn = torch.softmax(torch.einsum("bd,bpd->bp", q_cls, k) / 128.0, dim=-1)
        return xn, attn

    # ---------------------------------------------------------------- extract_bboxes_feats (bbox_utils.py:8-109), vectorised
    @torch.no_grad()
    def box_weights(self, boxes, g, gaussian: bool, variance: float):
        """[B,R,g*g] pooling weights of xywh pixel boxes: inclusive patch slices, uniform or separable Gaussian on
        linspace(-1, 1, span) (non-negative boxes; the reference's negative-index wrap is not reproduced here)."""
        b = torch.floor(boxes.to(self.dev).float() / PATCH)
        x1, y1 = b[..., 0], b[..., 1]
        x2, y2 = x1 + b[..., 2], y1 + b[..., 3]
        idx = torch.arange(g, device=self.dev, dtype=torch.float32)

        def axis(lo, hi):
            hi = torch.minimum(hi, torch.tensor(g - 1.0, device=self.dev))
            inside = (idx >= lo[..., None]) & (idx <= hi[..., None])
            span = (hi - lo + 1).clamp(min=1)[..., None]
            if gaussian:
                t = torch.where(span > 1, -1 + 2 * (idx - lo[..., None]) / (span - 1).clamp(min=1), torch.full_like(span, -1.0))
                f = torch.exp(-(t * t) / variance)
            else:
                f = torch.ones_like(inside, dtype=torch.float32)
            return f * inside

        wy, wx = axis(y1, y2), axis(x1, x2)                               # [B,R,g]
        w = wy[..., :, None] * wx[..., None, :]
        return (w / w.sum(dim=(-1, -2), keepdim=True)).flatten(2)

    @torch.no_grad()
    def pool(self, patch, weights):
        return torch.bmm(weights, patch)                                    # [B,R,P] x [B,P,768]

    # ---------------------------------------------------------------- Im2TxtProjector.project (im2txtprojection.py:353-385)
    @torch.no_grad()
    def project(self, q, chunk: int = 1024, temperature: float = 0.01):
        if self.bank is None:
            return q
        outs = []
        with self._ctx():
            for s in range(0, q.shape[0], chunk):                          # the reference chunks to bs * bs_factor rows too
                qq = q[s:s + chunk]
                qq = (qq / qq.norm(dim=-1, keepdim=True)).to(self.bank_n.dtype)
                p = torch.softmax((qq @ self.bank_n.T).float() / temperature, dim=-1)
                o = (p.to(self.bank.dtype) @ self.bank).float()
                outs.append(o / o.norm(dim=-1, keepdim=True))
        return torch.cat(outs, 0)

    # ---------------------------------------------------------------- decoding_batched (decap.py:116-160), GPT-2 4 x 4 x 192
    def _hidden(self, emb, kv, pos0):
        w, Tp = self.dw, "decoder.transformer."
        R, T, _ = emb.shape
        hd = D // N_HEAD
        x = emb + w[Tp + "wpe.weight"][pos0:pos0 + T]
        for i in range(N_LAYER):
            p = f"{Tp}h.{i}."
            h = F.layer_norm(x, (D,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], eps=1e-5)
            qkv = torch.addmm(w[p + "attn.c_attn.bias"], h.reshape(R * T, D), w[p + "attn.c_attn.weight"]).reshape(R, T, 3 * D)
            q, k, v = (t.reshape(R, T, N_HEAD, hd).transpose(1, 2) for t in qkv.split(D, dim=-1))
            if kv is not None:
                if kv[i] is not None:
                    k, v = torch.cat([kv[i][0], k], 2), torch.cat([kv[i][1], v], 2)
                kv[i] = (k, v)
            o = F.scaled_dot_product_attention(q, k, v, is_causal=(T > 1)).transpose(1, 2).reshape(R * T, D)
            x = x + torch.addmm(w[p + "attn.c_proj.bias"], o, w[p + "attn.c_proj.weight"]).reshape(R, T, D).float()
            h = F.layer_norm(x, (D,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], eps=1e-5).reshape(R * T, D)
            h = F.gelu(torch.addmm(w[p + "mlp.c_fc.bias"], h, w[p + "mlp.c_fc.weight"]), approximate="tanh")
            x = x + torch.addmm(w[p + "mlp.c_proj.bias"], h, w[p + "mlp.c_proj.weight"]).reshape(R, T, D).float()
        return F.layer_norm(x, (D,), w[Tp + "ln_f.weight"], w[Tp + "ln_f.bias"], eps=1e-5)

    @torch.no_grad()
    def decode(self, feats, steps: int = 30, use_cache: bool = False):
        w = self.dw
        wte = w["decoder.transformer.wte.weight"]
        with self._ctx():
            seq = F.linear(feats.float(), w["clip_project.model.0.weight"], w["clip_project.model.0.bias"]).float()[:, None]
            kv = [None] * N_LAYER if use_cache else None
            toks = []
            for t in range(steps):
                h = self._hidden(seq[:, -1:], kv, t)[:, -1] if use_cache else self._hidden(seq, None, 0)[:, -1]
                nxt = torch.argmax(torch.softmax(F.linear(h, wte).float(), -1), -1)   # argmax of the probabilities (:136,141)
                toks.append(nxt)
                seq = torch.cat([seq, wte[nxt][:, None]], 1)
        return torch.stack(toks, 1)

    # ---------------------------------------------------------------- one dense-captioning step (model.py:718-1041)
    @torch.no_grad()
    def dense_step(self, imgs, boxes, gaussian=True, variance=1.0, use_cache=False, chunk=1024):
        xn, _ = self.vit(imgs.to(self.dev))
        patch = xn[:, 1 + NREG:]
        g = int(math.isqrt(patch.shape[1]))
        feats = self.pool(patch, self.box_weights(boxes, g, gaussian, variance)).reshape(-1, D)
        ids = [self.decode(self.project(feats[s:s + chunk]), use_cache=use_cache) for s in range(0, feats.shape[0], chunk)]
        return torch.cat(ids, 0)
