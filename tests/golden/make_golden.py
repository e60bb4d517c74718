#!/usr/bin/env python
"""Generate tests/golden/*.pt by RUNNING THE UNMODIFIED REFERENCE (read-only /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference is imported as-is; the only interventions are import shims for packages that are
missing from this image and ``torch.hub.load`` (no network) being
pointed at a stand-in ``nn.Module`` with the attributes the reference touches
(``blocks[-1].attn.qkv``, ``.norm``, ``.patch_size``, ``forward(x, is_training=True) -> dict``).
Inputs come from the seeded generators in ``oracle/pipeline.py`` / ``oracle/*.make_weights``; the
fixtures store the REFERENCE'S outputs (and a checksum of each input so RNG drift is detected).
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/Patch-ioner"
sys.path.insert(0, ROOT)

from oracle import decap as o_decap  # noqa: E402
from oracle import dinov2 as o_vit  # noqa: E402
from oracle import pipeline as o_pipe  # noqa: E402


# ----------------------------------------------------------------------------- shims
def install_shims():
    import transformers

    if not hasattr(transformers, "AdamW"):
        transformers.AdamW = torch.optim.AdamW  # removed in transformers 5.x (decap.py:19)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, k):
            return _Any()

    for name in ["timm", "h5py", "open_clip", "loralib", "pycocotools", "fvcore"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub(name)
    if "ftfy" not in sys.modules:
        try:
            import ftfy  # noqa: F401
        except Exception:
            stub("ftfy", fix_text=lambda s: s)
    oc = sys.modules.get("open_clip")
    if oc is not None and not hasattr(oc, "utils"):
        u = stub("open_clip.utils", freeze_batch_norm_2d=lambda *a, **k: None)
        oc.utils = u

    return lambda: None


# ----------------------------------------------------------------------------- DINOv2 stand-in
class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (C // self.num_heads) ** -0.5, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((a @ v).transpose(1, 2).reshape(B, N, C))


class _LS(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return x * self.gamma


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(nn.functional.gelu(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, heads, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim, heads)
        self.ls1 = _LS(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)
        self.ls2 = _LS(dim)

    def forward(self, x):
        x = x + self.ls1(self.attn(self.norm1(x)))
        return x + self.ls2(self.mlp(self.norm2(x)))


class _PatchEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, 14, 14)


class HubStandIn(nn.Module):
    """Stand-in for torch.hub 'dinov2_vitb14_reg' (module form of oracle/dinov2.py)."""

    def __init__(self, depth=o_vit.DEPTH):
        super().__init__()
        D = o_vit.EMBED
        self.patch_size = 14
        self.cls_token = nn.Parameter(torch.zeros(1, 1, D))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + 37 * 37, D))
        self.register_tokens = nn.Parameter(torch.zeros(1, 4, D))
        self.mask_token = nn.Parameter(torch.zeros(1, D))
        self.patch_embed = _PatchEmbed(D)
        self.blocks = nn.ModuleList([_Block(D, o_vit.HEADS, o_vit.MLP) for _ in range(depth)])
        self.norm = nn.LayerNorm(D, eps=1e-6)

    def forward(self, x, is_training=False):
        B, _, H, W = x.shape
        g = H // 14
        x = self.patch_embed.proj(x).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(B, -1, -1), x], dim=1)
        x = x + o_vit.interpolate_pos_embed(self.pos_embed, g)
        x = torch.cat([x[:, :1], self.register_tokens.expand(B, -1, -1), x[:, 1:]], dim=1)
        for blk in self.blocks:
            x = blk(x)
        xn = self.norm(x)
        return {"x_norm_clstoken": xn[:, 0], "x_norm_regtokens": xn[:, 1:5],
                "x_norm_patchtokens": xn[:, 5:], "x_prenorm": x, "masks": None}


def csum(t: torch.Tensor) -> float:
    return float(t.double().sum())


def main():
    restore = install_shims()
    sys.path.insert(0, REF)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 8)
    import src.bbox_utils as ref_bbox
    import src.dino_extraction as ref_dex
    import src.embedding_utils as ref_emb
    import src.model as ref_model
    from src.decap import decap as ref_decap
    from src.decap.im2txtprojection.im2txtprojection import Im2TxtProjector

    # decap.py:67-69 unpickles a transformers-4.x GPT2Config, which lacks private fields that 5.x
    # expects; rebuild it from its public fields (same values) without touching the reference file.
    import pickle as _pickle
    from transformers import GPT2Config

    class _PickleShim:
        @staticmethod
        def load(f):
            c = _pickle.load(f)
            pub = {k: v for k, v in c.__dict__.items() if not k.startswith("_")}
            keep = ("vocab_size", "n_positions", "n_embd", "n_layer", "n_head", "n_inner", "activation_function",
                    "resid_pdrop", "embd_pdrop", "attn_pdrop", "layer_norm_epsilon", "initializer_range",
                    "scale_attn_weights", "use_cache", "scale_attn_by_inverse_layer_idx", "reorder_and_upcast_attn",
                    "bos_token_id", "eos_token_id", "tie_word_embeddings")
            return GPT2Config(**{k: pub[k] for k in keep if k in pub})

    ref_decap.pickle = _PickleShim

    out_dir = HERE
    meta = {"torch": torch.__version__}

    # ------------------------------------------------------------------ pooling (a5, a7)
    pool = {}
    for name, (B, g, R, D) in {"g16": (2, 16, 6, 64), "g37": (2, 37, 9, 64)}.items():
        S = g * 14
        gen = torch.Generator().manual_seed(100 + g)
        tok = torch.randn(B, g * g, D, generator=gen)
        amap = torch.rand(B, g * g, generator=gen).softmax(dim=-1)
        boxes = o_pipe.synth_boxes(B, R, S, seed=5 + g, degenerate_frac=0.2)
        boxes[0, 0] = torch.tensor([float(S - 20), float(S - 20), 100.0, 100.0])  # runs off the grid -> clamped
        boxes[1, 1] = torch.tensor([3.5, 7.25, 27.9, 13.99])                      # float floor-division
        boxes_set = boxes.clone()
        boxes_set[:, -1] = -1.0
        boxes_dense = boxes.clone()
        boxes_dense[:, -1] = torch.tensor([0.0, 0.0, 1.0, 1.0])
        rec = {"in_tok_sum": csum(tok), "in_amap_sum": csum(amap), "in_boxes": boxes.clone(),
               "shape": (B, g, R, D)}
        rec["mean"] = ref_bbox.extract_bboxes_feats(tok, boxes_dense.clone())
        rec["gauss_0.5"] = ref_bbox.extract_bboxes_feats(tok, boxes_dense.clone(), gaussian_avg=True, gaussian_bbox_variance=0.5)
        rec["gauss_1.0"] = ref_bbox.extract_bboxes_feats(tok, boxes_dense.clone(), gaussian_avg=True, gaussian_bbox_variance=1.0)
        a = amap.clone()
        rec["attn"] = ref_bbox.extract_bboxes_feats(tok, boxes_dense.clone(), attention_map=a)
        rec["attn_map_after"] = a.clone()  # the reference mutates it in place (Q2)
        rec["set_mean"] = ref_bbox.extract_bboxes_feats(tok, boxes_set.clone(), get_single_embedding_per_image=True)
        rec["set_gauss_1.0"] = ref_bbox.extract_bboxes_feats(tok, boxes_set.clone(), gaussian_avg=True, gaussian_bbox_variance=1.0,
                                                             get_single_embedding_per_image=True)
        rec["set_attn"] = ref_bbox.extract_bboxes_feats(tok, boxes_set.clone(), attention_map=amap.clone(),
                                                        get_single_embedding_per_image=True)
        # int boxes too (floor division on integers)
        rec["mean_intboxes"] = ref_bbox.extract_bboxes_feats(tok, boxes_dense.clone().long())
        # the integer 'pooling indices': patch-unit boxes exactly as the reference derives them
        bb = boxes_dense.clone()
        bb //= 14
        rec["patch_units"] = bb.int()
        rec["region_means_1"] = ref_model.compute_region_means(tok, 1)
        rec["region_means_100"] = ref_model.compute_region_means(tok, 100)
        rec["region_means_0.3"] = ref_model.compute_region_means(tok, 0.3)
        pool[name] = rec
    torch.save(pool, os.path.join(out_dir, "pooling.pt"))

    # ------------------------------------------------------------------ traces (a6)
    tr = {}
    for g in (16, 37):
        traces = o_pipe.synth_traces(4, seed=40 + g, n_min=20, n_max=80, outside_frac=0.1)
        # exact-boundary points: x = k/g exercises int(x / (1.0/g)) in double
        traces[0] += [{"x": k / g, "y": (g - k) / g, "t": 0.0} for k in range(g + 1)]
        traces[1] += [{"x": 1.0, "y": 1.0, "t": 0}, {"x": 0.0, "y": 0.0, "t": 0}, {"x": 0.29, "y": 0.57, "t": 0}]
        tr[f"g{g}"] = {"grids": torch.stack([ref_bbox.map_traces_to_grid(t, g) for t in traces]),
                       "npts": [len(t) for t in traces]}
        gen = torch.Generator().manual_seed(300 + g)
        tok = torch.randn(4, g * g, 64, generator=gen)
        sa = torch.rand(4, g * g, generator=gen).softmax(-1)
        rel = tr[f"g{g}"]["grids"]
        tr[f"g{g}"]["pool"] = (rel.unsqueeze(-1) * tok.view(4, g, g, 64)).mean(dim=(1, 2))        # model.py:1054
        rel2 = sa.view(rel.shape) * rel                                                           # model.py:1053
        tr[f"g{g}"]["pool_attn"] = (rel2.unsqueeze(-1) * tok.view(4, g, g, 64)).mean(dim=(1, 2))
        tr[f"g{g}"]["in_tok_sum"] = csum(tok)
    torch.save(tr, os.path.join(out_dir, "traces.pt"))

    # ------------------------------------------------------------------ CLS attention map (a3, a4)
    gen = torch.Generator().manual_seed(77)
    B, N, D = 2, 5 + 36, 768
    qkv = torch.randn(B, N, 3 * D, generator=gen)
    sa, maps = ref_dex.process_self_attention(qkv, B, N, 16, D, 0.125, 5, ret_self_attn_maps=True)
    patch = torch.randn(B, N - 5, D, generator=gen)
    torch.save({"in_qkv_sum": csum(qkv), "self_attn": sa.clone(), "self_attn_maps": maps.clone(),
                "avg_self_attn_token": (sa.unsqueeze(-1) * patch).mean(dim=1)},
               os.path.join(out_dir, "self_attn.pt"))

    # ------------------------------------------------------------------ memory projection (a8, a11)
    bank = o_pipe.synth_bank(3000, 768, seed=7, zero_frac=0.002)
    proj = object.__new__(Im2TxtProjector)
    proj.device = torch.device("cpu")
    proj.device_str = "cpu"
    emb = torch.tensor(bank.numpy())
    proj.embs_dataset = emb[emb.norm(dim=-1) != 0]  # im2txtprojection.py:343-345
    proj.text_dataset = None
    gen = torch.Generator().manual_seed(8)
    q = torch.randn(16, 768, generator=gen)
    q[3] = bank[11] * 2.5 + 0.01 * torch.randn(768, generator=gen)  # a near-duplicate of a bank row
    mem = {"in_q_sum": csum(q), "in_bank_sum": csum(bank), "M_after_filter": proj.embs_dataset.shape[0]}
    mem["out_norm"] = proj.project(q.clone(), normalize=True)
    mem["out_raw"] = proj.project(q.clone(), normalize=False)
    o, sims = proj.project(q.clone(), normalize=True, return_n_best_sims=5)
    mem["best_sims"] = torch.tensor(sims)
    t2d = torch.load(os.path.join(REF, "src/viecap/training/talk2dino/weights/vitb_mlp_infonce.pth"), map_location="cpu")
    A = t2d["linear_layer.weight"].float()
    b = t2d["linear_layer.bias"].float()
    A_pinv = ref_emb.get_pseudo_inverse(A)
    mem["revert"] = ref_emb.revert_transformation(mem["out_norm"], A_pinv=A_pinv, b=b)
    mem["A_pinv_sum"] = csum(A_pinv)
    torch.save(mem, os.path.join(out_dir, "memory.pt"))

    # ------------------------------------------------------------------ decoder (a9, a10)
    dec_w = o_decap.make_weights(seed=1234)
    ref_dec = ref_decap.DeCap(768)
    missing = ref_dec.load_state_dict(dec_w, strict=False)
    assert not [k for k in missing.missing_keys if "attn.bias" not in k and "masked_bias" not in k], missing
    ref_dec.eval()
    gen = torch.Generator().manual_seed(9)
    feats = torch.randn(6, 768, generator=gen)
    feats = feats / feats.norm(dim=-1, keepdim=True)
    rec_ids = []
    with torch.no_grad():
        ref_decap.decoding_batched(ref_dec, feats, decoding_method=lambda ids: (rec_ids.append([int(i) for i in ids]) or ""))
        _, scores = ref_decap.decoding_batched(ref_dec, feats, compute_scores=True, decoding_method=lambda ids: "")
        # logits of the first step, to pin the GPT-2 arithmetic itself
        e0 = ref_dec.clip_project(feats).view(6, 1, -1)
        logits0 = ref_dec.decoder(inputs_embeds=e0).logits[:, -1]
    torch.save({"in_feats_sum": csum(feats), "ids": torch.tensor(rec_ids), "scores": torch.tensor(scores),
                "logits0_top_values": logits0.topk(8, dim=-1).values.clone(), "logits0_top_indices": logits0.topk(8, dim=-1).indices.clone(), "wte_sum": csum(dec_w["decoder.transformer.wte.weight"])},
               os.path.join(out_dir, "decoder.pt"))

    # ------------------------------------------------------------------ full Patchioner.forward (a12)
    vit_w = o_vit.make_weights(seed=1234)
    stand_in = HubStandIn()
    sd = {k: v for k, v in vit_w.items()}
    stand_in.load_state_dict(sd, strict=True)
    real_hub_load = torch.hub.load
    torch.hub.load = lambda *a, **k: stand_in
    tmp = tempfile.NamedTemporaryFile(suffix=".pt", delete=False)
    torch.save(dec_w, tmp.name)
    full = {}
    try:
        for variant, with_bank in (("decap", True), ("capdec", False)):
            model = ref_model.Patchioner.from_config({
                "decap_weights": tmp.name, "prefix_size": 768, "linear_talk2dino": False,
                "support_memory_size": 0, "dino_model": "dinov2_vitb14_reg", "normalize": True,
                "resize_dim": 224, "crop_dim": 224}, device="cpu")
            if with_bank:
                model.im_proj = proj
            ids_rec = []
            model.decoding_method = lambda ids: (ids_rec.append([int(i) for i in ids]) or "")
            B, S, R = 2, 224, 4
            imgs = o_pipe.synth_images(B, S, seed=1)
            boxes = o_pipe.synth_boxes(B, R, S, seed=1, pad="dense")
            boxes_set = o_pipe.synth_boxes(B, R, S, seed=2, pad="set")
            traces = o_pipe.synth_traces(B, seed=1)
            rec = {"in_img_sum": csum(imgs)}

            def run(key, **kw):
                ids_rec.clear()
                with torch.no_grad():
                    model(imgs, **kw)
                rec[key] = torch.tensor(ids_rec)

            run("cls+bbox_mean", get_cls_capt=True, bboxes=boxes.clone())
            run("bbox_gauss1", get_cls_capt=False, bboxes=boxes.clone(), gaussian_avg=True, gaussian_bbox_variance=1.0)
            run("bbox_attn", get_cls_capt=False, bboxes=boxes.clone(), use_attn_map_for_bboxes=True)
            run("set_gauss1", get_cls_capt=False, bboxes=boxes_set.clone(), get_controllable_capts=True,
                gaussian_avg=True, gaussian_bbox_variance=1.0)
            run("trace", get_cls_capt=False, traces=traces)
            run("trace_attn", get_cls_capt=False, traces=traces, use_attention_tracing=True)
            run("avg_self_attn+avg_patch", get_cls_capt=False, get_avg_self_attn_capt=True, get_avg_patch_capt=True,
                gaussian_img_variance=1.0)
            full[variant] = rec
        # the ViT outputs the stand-in produced (module form) -- pins oracle/dinov2.forward (functional form)
        with torch.no_grad():
            d = stand_in(imgs, is_training=True)
        full["vit_cls"] = d["x_norm_clstoken"].clone()
        full["vit_patch_0_100"] = d["x_norm_patchtokens"][:, :100].clone()
    finally:
        torch.hub.load = real_hub_load
        os.unlink(tmp.name)
    torch.save(full, os.path.join(out_dir, "forward.pt"))
    torch.save(meta, os.path.join(out_dir, "meta.pt"))
    restore()
    for f in sorted(os.listdir(out_dir)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(out_dir, f)))


if __name__ == "__main__":
    main()
