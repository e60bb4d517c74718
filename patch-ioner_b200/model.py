"""``Patchioner`` -- drop-in for the reference's model facade on the talk2dino (DINOv2 ViT-B/14-reg) path.

Mirrors Patch-ioner/src/model.py: ``Patchioner.from_config`` (:666-715), ``forward`` (:718-1058, same keyword
names, same output keys), ``caption_tokens`` (:1392-1423).  Underneath, every numeric step runs in
libpio_sm100.so (hand-written sm_100a CUDA) -- there is no PyTorch compute path and no CPU fallback.

Scope: DINOv2-reg backbone + DeCap / CapDec text side, plus the ViECap captioner on the same
embeddings (``viecap:`` config key -> ``viecap.py``).  The alternative backbones and captioners of the
reference (ProxyCLIP, RegionCLIP, INViTE, DenseCLIP, AlphaCLIP, MeaCap, ClipCap) are out of scope and raise
``NotImplementedError`` instead of silently doing something else.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib as L
from . import ops
from .detok import default_batch_detokenizer

EOT = "<|endoftext|>"
SOT = "<|startoftext|>"


def _load_tensor_file(path: str):
    if path.endswith((".pt", ".pth", ".bin")):
        return torch.load(path, map_location="cpu", weights_only=False)
    if path.endswith(".npy"):
        import numpy as np

        return torch.from_numpy(np.load(path))
    if path.endswith((".h5", ".hdf5")):
        try:
            import h5py  # the reference's bank format (im2txtprojection.py:320-323)
        except ImportError as e:  # pragma: no cover - h5py is absent from this image
            raise RuntimeError("reading an HDF5 memory bank needs h5py") from e
        with h5py.File(path, "r") as hf:
            key = [k for k in hf.keys() if k.endswith("-embeddings")][0]
            return torch.from_numpy(hf[key][:])
    raise ValueError(f"do not know how to read {path}")


def _load_h5_texts(path: str):
    try:
        import h5py
    except ImportError:  # pragma: no cover - h5py is absent from this image
        return None
    with h5py.File(path, "r") as hf:
        keys = [k for k in hf.keys() if k.endswith("-text")]
        return list(hf[keys[0]][:]) if keys else None


class Patchioner:
    """See module docstring.  Not an ``nn.Module``: the weights live in library-owned device memory."""

    def __init__(self, decoder_weights, device, prefix_size, linear_talk2dino=False, support_memory_size=0,
                 projection_type=None, dino_model=None, proxyclip_clipmodel=None, proxyclip_vfm=None,
                 use_talk2dino_project=True, normalize=True, attention_type="qkv", talk2dino_config=None,
                 talk2dino_weights=None, resize_dim=518, crop_dim=518, talk2dino_attn_type="qkv",
                 calculate_argmax_text=False, online_texts=None, clip_model_name=None, use_open_clip=False,
                 viecap_config=None, regionclip_config=None, invite_config=None, denseclip_config=None,
                 alphaclip_config=None, clipcap_config=None, hf_repo_id=None,
                 # extensions (no network in this image: weights are given as files / state dicts)
                 dino_weights=None, memory_bank=None, memory_bank_texts=None, precision="fp32", **kwargs):
        if viecap_config is not None and viecap_config.get("meacap", False):
            raise NotImplementedError("viecap.meacap selects MeaCap (model.py:108-110): outside the B200 hot path")
        for name, val in (("proxyclip_clipmodel", proxyclip_clipmodel),
                          ("regionclip_config", regionclip_config), ("invite_config", invite_config),
                          ("denseclip_config", denseclip_config), ("alphaclip_config", alphaclip_config),
                          ("clipcap", clipcap_config)):
            if val is not None:
                raise NotImplementedError(f"'{name}' selects a backbone/captioner outside the B200 hot path (SURVEY.md 8)")
        if online_texts is not None:
            raise NotImplementedError("online_texts builds a bank with a CLIP text encoder (im2txtprojection.py:448-560): out of scope")
        h5_bank = isinstance(memory_bank, str) and memory_bank.endswith((".h5", ".hdf5", ".h5.pt"))
        if calculate_argmax_text and ((memory_bank_texts is None and not h5_bank) or support_memory_size <= 0):
            raise ValueError("calculate_argmax_text needs the bank and its captions (memory_bank_texts, or an HDF5 bank with a '-text' dataset)")
        if dino_model is None or "dinov2" not in dino_model or "vitb14" not in dino_model or "reg" not in dino_model:
            raise NotImplementedError(f"dino_model={dino_model!r}: only 'dinov2_vitb14_reg' is built (model.py:342-343)")
        if attention_type != "qkv":
            raise NotImplementedError("attention_type != 'qkv' re-wires the last block (model.py:557-582): out of scope")
        if precision not in ops.MODES:
            raise ValueError(f"precision must be one of {list(ops.MODES)}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PioError("patchioner_b200 runs on CUDA (sm_100a) only: there is no CPU fallback")
        L.lib()  # fail loudly now if the CUDA library has not been built
        self.precision = precision
        self.decoding_method: Optional[Callable[[List[int]], str]] = None  # model.py:105
        self.normalize = normalize
        self.resize_dim, self.crop_dim = resize_dim, crop_dim
        self.model_name = dino_model
        self.num_global_tokens = 5
        self.patch_size = 14
        self.embed_dim = 768
        self.num_attn_heads = 16  # sic (model.py:336); only get_attn_heads_capt depends on it
        self.num_patch_tokens = crop_dim // 14 * crop_dim // 14
        self.num_tokens = self.num_global_tokens + self.num_patch_tokens
        self.backbone_type = "DINO"
        self.viecap = self.clipcap = None
        self.calculate_argmax_text = bool(calculate_argmax_text)
        # indexed by the row of the zero-row-FILTERED bank, like the reference (text_dataset is not filtered, embs are:
        # im2txtprojection.py:342-345, :372)
        self.text_dataset = list(memory_bank_texts) if memory_bank_texts is not None else None

        # --- backbone (model.py:342-343 loads it from torch.hub; here: a state dict with the hub's key names)
        if isinstance(dino_weights, str):
            dino_weights = torch.load(dino_weights, map_location="cpu", weights_only=False)
        if dino_weights is None:
            raise ValueError("dino_weights (state dict or path) is required: torch.hub needs the network")
        self.dino = ops.Vit(dino_weights, self.device, precision)

        # --- decoder (model.py:165-166 -> decap.py:188-222)
        if isinstance(decoder_weights, str) and viecap_config is not None and not os.path.exists(decoder_weights):
            decoder_weights = None  # the ViECap configs still name a DeCap checkpoint, which caption_tokens never uses
        if isinstance(decoder_weights, str):
            decoder_weights = torch.load(decoder_weights, map_location="cpu", weights_only=False)
        if viecap_config is not None:
            # model.py:107-113: the ViECap captioner replaces the DeCap decode in caption_tokens (model.py:1394-1398)
            from .viecap import VieCap
            self.viecap = VieCap(viecap_config, self.device, clip_model_name, precision=precision)
        self.decoder = None
        if decoder_weights is None and self.viecap is None:
            raise ValueError("decap_weights is required")
        if decoder_weights is not None:
            self.decoder = ops.Decoder(decoder_weights, self.device, precision)
            if self.decoder.prefix_size != prefix_size:
                raise ValueError(f"prefix_size {prefix_size} != clip_project input {self.decoder.prefix_size}")

        # --- caption memory (model.py:144-186).  support_memory_size == 0 -> CapDec, no bank.
        self.im_proj = None
        if support_memory_size > 0:
            bank = memory_bank
            if bank is None and isinstance(projection_type, str) and os.path.exists(projection_type):
                bank = projection_type
            if isinstance(bank, str):
                if memory_bank_texts is None and bank.endswith((".h5", ".hdf5")):
                    memory_bank_texts = _load_h5_texts(bank)  # the '{name}-text' dataset next to the embeddings (:320-323)
                    self.text_dataset = memory_bank_texts
                bank = _load_tensor_file(bank)
                if isinstance(bank, dict):  # bank_builder.write_bank's torch flavour: '{name}-embeddings' (+ '{name}-text')
                    tk = [k for k in bank if k.endswith("-text")]
                    if memory_bank_texts is None and tk:
                        self.text_dataset = memory_bank_texts = list(bank[tk[0]])
                    bank = bank[[k for k in bank if k.endswith("-embeddings")][0]]
            if bank is None:
                raise ValueError("support_memory_size > 0 needs memory_bank (tensor or file); building banks from "
                                 "captions needs CLIP text encoders + network (out of scope)")
            self.im_proj = ops.Bank(bank, self.device, precision)
        if self.calculate_argmax_text and self.text_dataset is None:
            raise ValueError("calculate_argmax_text: the memory bank file carries no captions")

        # --- Talk2DINO inversion (model.py:618-627)
        self.embed_inversion = False
        if talk2dino_weights is not None:
            from .talk2dino import pseudo_inverse

            sd = torch.load(talk2dino_weights, map_location="cpu", weights_only=False) if isinstance(talk2dino_weights, str) else talk2dino_weights
            A = sd["linear_layer.weight"].float()
            b = sd["linear_layer.bias"].float()
            A_pinv = pseudo_inverse(A)                      # init-time SVD on the host (embedding_utils.py:3-15)
            self.talk2dino_A_pinv = A_pinv.to(self.device).contiguous()          # [512, 768]
            self.talk2dino_b = b.to(self.device)
            # (x - b) @ A_pinv^T  ==  x @ A_pinv^T - A_pinv b   -> one dense layer with a folded bias
            self._inv_bias = (-(A_pinv @ b)).to(self.device).contiguous()
            self._inv_w = self.talk2dino_A_pinv.to(torch.bfloat16) if precision == "bf16" else self.talk2dino_A_pinv
            self.embed_inversion = True

        self._transforms = None

    # ------------------------------------------------------------------------------------------ config
    @classmethod
    def from_config(cls, config, device="cuda", online_texts=None, **overrides):
        """model.py:666-715: ``config`` is a dict or a YAML path with the reference's keys."""
        if isinstance(config, str):
            if not os.path.exists(config):
                raise FileNotFoundError(f"{config}: HuggingFace repo ids need the network (hf_utils.py) -- not available")
            import yaml

            with open(config, "r") as f:
                config = yaml.safe_load(f)
        config = dict(config)
        config.update(overrides)
        return cls(
            projection_type=config.get("projection_type", "coco"),
            decoder_weights=config.get("decap_weights", None),
            device=device,
            prefix_size=config["prefix_size"],
            linear_talk2dino=config.get("linear_talk2dino", False),
            support_memory_size=config["support_memory_size"],
            dino_model=config.get("dino_model", None),
            proxyclip_clipmodel=config.get("proxyclip_clipmodel", None),
            proxyclip_vfm=config.get("proxyclip_vfm", None),
            use_talk2dino_project=config.get("use_talk2dino_project", True),
            normalize=config.get("normalize", True),
            attention_type=config.get("attention_type", "qkv"),
            talk2dino_config=config.get("talk2dino_config", None),
            talk2dino_weights=config.get("talk2dino_weights", None),
            resize_dim=config.get("resize_dim", 518),
            crop_dim=config.get("crop_dim", 518),
            talk2dino_attn_type=config.get("talk2dino_attn_type", "qkv"),
            calculate_argmax_text=config.get("calculate_argmax_text", False),
            clip_model_name=config.get("clip_model_name", None),
            online_texts=online_texts,
            use_open_clip=config.get("use_open_clip", False),
            viecap_config=config.get("viecap", None),
            regionclip_config=config.get("regionclip_config", None),
            invite_config=config.get("invite_config", None),
            denseclip_config=config.get("denseclip_config", None),
            alphaclip_config=config.get("alphaclip_config", None),
            clipcap_config=config.get("clipcap", None),
            hf_repo_id=config.get("hf_repo_id", None),
            dino_weights=config.get("dino_weights", None),
            memory_bank=config.get("memory_bank", None),
            memory_bank_texts=config.get("memory_bank_texts", None),
            precision=config.get("precision", "fp32"),
        )

    # image transforms the eval drivers read (model.py:347-357); built lazily (torchvision is host-side only)
    def _build_transforms(self):
        import torchvision.transforms as T

        norm = T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))
        self._transforms = (
            T.Compose([T.Resize(self.resize_dim, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(self.crop_dim),
                       T.ToTensor(), norm]),
            T.Compose([T.Resize((self.resize_dim, self.resize_dim), interpolation=T.InterpolationMode.BICUBIC),
                       T.ToTensor(), norm]))

    @property
    def image_transforms(self):
        if self._transforms is None:
            self._build_transforms()
        return self._transforms[0]

    @property
    def image_transforms_no_crop(self):
        if self._transforms is None:
            self._build_transforms()
        return self._transforms[1]

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device and torch.device(device).type != self.device.type:
            raise L.PioError("weights live in library-owned memory on the construction device")
        return self

    def parameters(self):
        yield torch.empty(0, device=self.device)

    # ------------------------------------------------------------------------------------------ text side
    def embed_tokens(self, dino_tokens: torch.Tensor, project: bool = True) -> torch.Tensor:
        """model.py:1406-1422 up to the decoder input: memory projection iff a bank exists, optional inversion."""
        x = dino_tokens
        if self.im_proj is not None and project:
            x = self.im_proj.project(x, normalize=self.normalize)
        if self.embed_inversion:
            a = x.to(torch.bfloat16) if self.precision == "bf16" else x.float().contiguous()
            x = ops.linear(a, self._inv_w, self.precision, bias=self._inv_bias)
        return x

    def _project_rows(self, x: torch.Tensor, chunk: int = 8192) -> torch.Tensor:
        """im_proj.project(x, normalize=True) for many rows (chunked to bound the workspace)."""
        x = x.reshape(-1, x.shape[-1])
        return torch.cat([self.im_proj.project(x[s:s + chunk], normalize=True) for s in range(0, x.shape[0], chunk)], 0)

    def caption_token_ids(self, dino_tokens: torch.Tensor, project: bool = True, compute_scores: bool = False):
        """ids int32 [R,30] on the device (+ scores).  Chunked only to bound the decoder workspace."""
        feats = dino_tokens.reshape(-1, dino_tokens.shape[-1])
        ids, scores = [], []
        chunk = 8192
        for s in range(0, max(feats.shape[0], 1), chunk):
            f = feats[s:s + chunk]
            if f.shape[0] == 0:
                break
            pre = self.embed_tokens(f, project)
            r = self.decoder.decode(pre, 30, compute_scores)
            if compute_scores:
                ids.append(r[0])
                scores.append(torch.exp(r[1]))  # decap.py:159-160
            else:
                ids.append(r)
        ids = torch.cat(ids, 0) if ids else torch.empty(0, 30, dtype=torch.int32, device=self.device)
        return (ids, torch.cat(scores, 0)) if compute_scores else ids

    def _ids_to_text(self, ids: torch.Tensor) -> List[str]:
        """decap.py:162-181: detokenise, cut at <|endoftext|>."""
        return self._host_ids_to_text(ids.cpu())  # the one device->host read of a caption batch

    def _host_ids_to_text(self, host: torch.Tensor) -> List[str]:
        """Default: ONE batched call (detok.BatchDetokenizer -> pio_detok_rows) for the whole id matrix; a user
        ``decoding_method`` hook (model.py:105) is called row by row like the reference does."""
        if self.decoding_method is None:
            return default_batch_detokenizer()(host)
        return [self.decoding_method(r).split(EOT)[0].replace(SOT, "") for r in host.tolist()]

    def caption_tokens(self, dino_tokens, project=True, return_n_best_sims=None, compute_scores: bool = False, rows_per_call=None):
        """model.py:1392-1423.  ``rows_per_call`` (extension): the rows are the concatenation of reference calls of that many
        rows each -- only ViECap's output depends on it (hard prompts are padded per call)."""
        if self.viecap is not None:  # model.py:1394-1398
            if return_n_best_sims:
                raise Exception("return_n_best_sims is not supported with viecap")
            return self.viecap.forward(dino_tokens.reshape(-1, dino_tokens.shape[-1]), compute_scores=compute_scores,
                                       pad_group=rows_per_call)
        if self.calculate_argmax_text:
            # model.py:1408-1411 -> im2txtprojection.py:371-375: the caption is the text of the most similar bank row
            feats = dino_tokens.reshape(-1, dino_tokens.shape[-1])
            sims, rows = self.im_proj.best_sims(feats, int(return_n_best_sims or 1), with_rows=True)
            texts = [self.text_dataset[i] for i in rows[:, 0].cpu().tolist()]
            captions = [t.decode() if isinstance(t, (bytes, bytearray)) else t for t in texts]
            ret = (captions, sims.cpu().tolist()) if return_n_best_sims else captions
            return (ret, [1.0] * len(captions)) if compute_scores else ret
        if return_n_best_sims:
            # the reference only honours it together with calculate_argmax_text (model.py:1408-1411); with the decoder
            # path its forward() fails while unpacking (model.py:1033)
            raise ValueError("return_n_best_sims needs calculate_argmax_text=True (as in the reference, model.py:1408-1411)")
        r = self.caption_token_ids(dino_tokens, project, compute_scores)
        if compute_scores:
            return self._ids_to_text(r[0]), r[1].cpu().tolist()
        return self._ids_to_text(r)

    # ------------------------------------------------------------------------------------------ forward
    def region_embeddings(self, imgs, bboxes=None, traces=None, masks=None, get_controllable_capts=False, gaussian_avg=False,
                          gaussian_bbox_variance=0.5, use_attn_map_for_bboxes=False, use_attention_tracing=False):
        """ViT + pooling only: dict of region embeddings on the device (what ``forward`` captions)."""
        tokens, attn, _ = self.dino.forward(imgs, want_attn=True)
        patch = tokens[:, self.num_global_tokens:]
        out: Dict[str, torch.Tensor] = {"cls": tokens[:, 0], "registers": tokens[:, 1:5], "patch": patch, "self_attn": attn}
        bs, P, _ = patch.shape
        g = int(P ** 0.5)
        if bboxes is not None:
            amap = attn if use_attn_map_for_bboxes else None
            out["set" if get_controllable_capts else "bbox"] = ops.pool_boxes(
                patch, bboxes, self.patch_size, gaussian_avg, gaussian_bbox_variance, amap,
                get_single_embedding_per_image=get_controllable_capts)
        if traces is not None:
            w = ops.trace_bins(traces, g, patch.device, attn if use_attention_tracing else None)
            out["trace"] = ops.pool_grid(patch, w.reshape(bs, 1, P), 1.0 / P)[:, 0]       # model.py:1054 (.mean over g*g)
        if masks is not None:
            m = masks.to(patch.device, torch.float32)
            out["mask"] = ops.pool_grid(patch, m.reshape(bs, -1, P), 1.0 / P)
        return out

    @torch.no_grad()
    def forward(self, imgs, get_cls_capt=True, get_avg_self_attn_capt=False, get_attn_heads_capt=False, get_patch_capts=False,
                get_register_capts=False, bboxes=None, traces=None, get_controllable_capts=False, bs_factor=4,
                gaussian_avg=False, gaussian_bbox_variance=0.5, get_avg_patch_capt=False, gaussian_img_variance=1,
                use_attn_map_for_bboxes=False, use_attention_tracing=False, double_DINO_for_bboxes=False,
                double_DINO_for_bboxes_return_type="avg", double_DINO_use_cls=False, cleaning_type=None,
                clean_after_projection=True, alpha=1.0, clean_from="cls", caption_bboxes_type: str = None,
                return_n_best_sims=None, compute_scores: bool = False, masks=None, return_ids: bool = False):
        """Same keywords and output keys as the reference (model.py:718-1058).  Extensions: ``masks`` [B,R,g,g]
         -> ``mask_capts``; ``return_ids=True`` returns int32 id tensors instead of strings."""
        assert clean_from in ["cls", "avg_self_attn"]
        assert cleaning_type in [None, "orthogonal_projection", "contrastive_mask"]
        if caption_bboxes_type is not None:  # model.py:770-771: caption the CROP of every box as a whole image
            return self.caption_bboxes(imgs, bboxes, caption_bboxes_type, compute_scores=compute_scores)
        if double_DINO_for_bboxes and (bboxes is None or get_controllable_capts):
            raise ValueError("double_DINO_for_bboxes applies to the per-box branch (bboxes given, get_controllable_capts=False)")
        if cleaning_type is not None and self.im_proj is None:
            raise ValueError("cleaning_type needs the caption memory (the reference calls im_proj.project, model.py:895-913)")
        if self.calculate_argmax_text and return_ids:
            raise ValueError("return_ids has no meaning with calculate_argmax_text (captions are bank texts)")
        imgs = imgs.to(self.device, non_blocking=True)
        outs: Dict[str, object] = {}
        bs = imgs.shape[0]
        tokens, self_attn, qkv_last = self.dino.forward(imgs, want_attn=True, want_qkv=bool(get_attn_heads_capt))
        cls, reg, patch = tokens[:, 0], tokens[:, 1:5], tokens[:, 5:]
        P, D = patch.shape[1], patch.shape[2]
        g = int(P ** 0.5)

        avg_self_attn_token = None
        if get_avg_self_attn_capt or (cleaning_type is not None and clean_from == "avg_self_attn"):
            avg_self_attn_token = ops.pool_grid(patch, self_attn.reshape(bs, 1, P), 1.0 / P)[:, 0]  # model.py:869, before any cleaning
        project_regions = True
        patch_orig = patch  # the per-head tokens are taken before any cleaning (model.py:872)
        if cleaning_type is not None:
            # model.py:879-922: the patch tokens are REPLACED by context-cleaned, memory-projected tokens; the patch and box
            # captions then skip the projection (project = cleaning_type is None, model.py:966, 1014)
            ctx = cls if clean_from == "cls" else avg_self_attn_token
            if clean_after_projection:
                pp = self._project_rows(patch.reshape(-1, D)).reshape(bs, P, D)
                patch = ops.ctx_clean(pp, self._project_rows(ctx), cleaning_type, alpha)
            else:
                cleaned = ops.ctx_clean(patch, ctx.contiguous(), cleaning_type, alpha, prenorm=True)
                patch = self._project_rows(cleaned.reshape(-1, D)).reshape(bs, P, D)
            project_regions = False

        def emit_texts(key, feats, group=None):
            """calculate_argmax_text (model.py:1408-1411): captions are bank texts; only the box branch asks for sims."""
            want_sims = return_n_best_sims if key == "bbox_capts" else None
            # the reference captions boxes in calls of bs * bs_factor regions (model.py:981-1013), everything else in one call
            per_call = bs * bs_factor if key == "bbox_capts" else None
            ret = self.caption_tokens(feats, return_n_best_sims=want_sims, compute_scores=compute_scores, rows_per_call=per_call)
            ret, sc = (ret if compute_scores else (ret, None))
            capts, sims = (ret if want_sims else (ret, None))
            cut = (lambda v: [v[i * group:(i + 1) * group] for i in range(bs)]) if group is not None else (lambda v: v)
            outs[key] = cut(capts)
            if sims is not None:
                outs["bbox_sims"] = cut(sims)
            if compute_scores:
                skey = {"bbox_capts": "bbox_scores", "patch_tokens_capts": "patch_tokens_scores", "register_capts": "register_scores",
                        "attn_heads_capts": "attn_heads_scores"}.get(key, key + "_scores")
                outs[skey] = cut(sc)

        def emit(key, feats, group=None, project=True):
            if feats.shape[0] == 0:  # e.g. bboxes of shape [B, 0, 4]: no region, no caption
                outs[key] = (torch.empty(bs, 0, 30, dtype=torch.int32, device=self.device) if return_ids else [[] for _ in range(bs)])
                if compute_scores:
                    outs[{"bbox_capts": "bbox_scores"}.get(key, key + "_scores")] = [[] for _ in range(bs)]
                return None
            if self.viecap is not None and return_ids:  # extension: the 64 generated ids per region, before the sentence cut
                if return_n_best_sims is not None and key == "bbox_capts":
                    raise Exception("return_n_best_sims is not supported with viecap")
                ids = self.viecap.forward_ids(feats, pad_group=bs * bs_factor if key == "bbox_capts" else None)
                outs[key] = ids.reshape(bs, group, -1) if group is not None else ids
                return None
            if self.calculate_argmax_text or self.viecap is not None:
                return emit_texts(key, feats, group)
            if return_n_best_sims is not None and key == "bbox_capts":
                self.caption_tokens(feats[:0], return_n_best_sims=return_n_best_sims)  # raises like the reference's decoder path
            r = self.caption_token_ids(feats, project=project, compute_scores=compute_scores)
            ids, sc = (r if compute_scores else (r, None))
            vals = ids if return_ids else self._ids_to_text(ids)
            if group is not None:
                vals = ids.reshape(bs, group, -1) if return_ids else [vals[i * group:(i + 1) * group] for i in range(bs)]
            outs[key] = vals
            if compute_scores:
                s = sc.cpu().tolist()
                skey = {"bbox_capts": "bbox_scores", "patch_tokens_capts": "patch_tokens_scores",
                        "register_capts": "register_scores", "attn_heads_capts": "attn_heads_scores"}.get(key, key + "_scores")
                outs[skey] = s if group is None else [s[i * group:(i + 1) * group] for i in range(bs)]

        if get_cls_capt:
            emit("cls_capt", cls)
        if get_avg_self_attn_capt:  # model.py:869
            emit("avg_self_attn_capt", avg_self_attn_token)
        if get_avg_patch_capt:      # model.py:45-94
            if gaussian_img_variance == 0:  # model.py:71-79: one-hot at a (python-random, for an even grid) central patch
                emit("avg_patch_capt", ops.region_centre_rows(patch))
            else:
                w = ops.region_mean_weights(g, gaussian_img_variance, patch.device)
                emit("avg_patch_capt", ops.pool_grid(patch, w.reshape(1, 1, P).expand(bs, 1, P), 1.0)[:, 0])
        if get_attn_heads_capt:     # model.py:871-872, 950-960: one embedding per "head" map (16 x 48-channel re-cut, sic)
            maps = ops.cls_head_attention(qkv_last, self.num_global_tokens, self.num_attn_heads, 0.125)
            emit("attn_heads_capts", ops.pool_grid(patch_orig, maps, 1.0 / P).reshape(-1, D), group=self.num_attn_heads)
        if get_patch_capts:
            emit("patch_tokens_capts", patch.reshape(-1, D), group=P, project=project_regions)
        if get_register_capts:
            emit("register_capts", reg.reshape(-1, D), group=4)
        if bboxes is not None and not get_controllable_capts and double_DINO_for_bboxes:
            # model.py:983-992: the hooked output of blocks[-1], normalised, is exactly `tokens`; note that the reference hands over
            # the (possibly cleaned) patch tokens' source, i.e. the un-cleaned layer output
            feats = self.double_dino_feats(tokens, bboxes, double_DINO_for_bboxes_return_type, double_DINO_use_cls, gaussian_bbox_variance)
            emit("bbox_capts", feats.reshape(-1, D), group=bboxes.shape[1], project=project_regions)
        elif bboxes is not None and not get_controllable_capts:
            amap = self_attn if use_attn_map_for_bboxes else None
            feats = ops.pool_boxes(patch, bboxes, self.patch_size, gaussian_avg, gaussian_bbox_variance, amap)
            emit("bbox_capts", feats.reshape(-1, D), group=bboxes.shape[1], project=project_regions)
        elif bboxes is not None and get_controllable_capts:
            amap = self_attn if use_attn_map_for_bboxes else None
            feats = ops.pool_boxes(patch, bboxes, self.patch_size, gaussian_avg, gaussian_bbox_variance, amap,
                                   get_single_embedding_per_image=True)
            emit("set_controllable_capts", feats)
        if traces is not None:
            w = ops.trace_bins(traces, g, patch.device, self_attn if use_attention_tracing else None)
            emit("trace_capts", ops.pool_grid(patch, w.reshape(bs, 1, P), 1.0 / P)[:, 0])
        if masks is not None:
            m = masks.to(patch.device, torch.float32).reshape(bs, -1, P)
            emit("mask_capts", ops.pool_grid(patch, m, 1.0 / P).reshape(-1, D), group=m.shape[1])
        return outs

    __call__ = forward

    def double_dino_feats(self, tokens: torch.Tensor, bboxes: torch.Tensor, return_type: str = "avg", use_cls: bool = False,
                          gaussian_bbox_variance: float = 0.5) -> torch.Tensor:
        """extract_bboxes_feats_double_dino (bbox_utils.py:300-403): for every box the last block runs again on
        [cls | 4 registers | the box's patch tokens] (use_cls) or on the patch tokens alone; 'avg' = mean of the patch outputs,
        'cls' = output of the cls position, 'gaussian_avg' = Gaussian pooling of the block's INPUT patches (sic, :377-389).
        Quirk kept: the box is floor-divided by the patch size and then sliced as [y : h + 1, x : w + 1] -- width and height are
        used as END indices (:329) -- with Python slice clamping.  -> [B, R, 768]."""
        if return_type not in ("avg", "cls", "gaussian_avg"):
            raise ValueError(f"double_DINO return type {return_type!r}")
        if return_type == "cls" and not use_cls:
            raise AssertionError("return_type 'cls' needs double_DINO_use_cls (bbox_utils.py:340)")
        B, N, D = tokens.shape
        P = N - self.num_global_tokens
        g = int(P ** 0.5)
        bb = bboxes.detach().clone().cpu()
        bb //= self.patch_size        # torch floor division, like :319-320
        bb = bb.int().tolist()
        n_glob = self.num_global_tokens if use_cls else 0
        seqs = []                      # (length, b, j, token indices)
        for b in range(B):
            for j, (x, y, w, h) in enumerate(bb[b]):
                ys, xs = range(g)[y:h + 1], range(g)[x:w + 1]      # Python slice semantics (negative wrap, clamping)
                idx = [b * N + t for t in range(n_glob)] + [b * N + self.num_global_tokens + yy * g + xx for yy in ys for xx in xs]
                seqs.append((len(idx), b, j, idx, (ys[0] if len(ys) else 0, xs[0] if len(xs) else 0, len(ys), len(xs))))
        R = bboxes.shape[1]
        if return_type == "gaussian_avg":
            # Gaussian pooling of the block's INPUT patches over the sliced rectangle (:377-389; the second pass through the block
            # does not enter the result): the standard box kernel in patch units (patch_size 1 -> span = w + 1)
            rect = torch.tensor([[[float(r[1]), float(r[0]), float(max(r[3] - 1, 0)), float(max(r[2] - 1, 0))]
                                  for (_, _, _, _, r) in seqs[b * R:(b + 1) * R]] for b in range(B)])
            empty = torch.tensor([[r[2] == 0 or r[3] == 0 for (_, _, _, _, r) in seqs[b * R:(b + 1) * R]] for b in range(B)])
            out = ops.pool_boxes(tokens[:, self.num_global_tokens:], rect, 1, True, gaussian_bbox_variance)
            if empty.any():  # an empty rectangle sums to zeros in the reference
                out[empty.to(out.device)] = 0.0
            return out
        order = sorted(range(len(seqs)), key=lambda i: seqs[i][0])
        lens = [seqs[i][0] for i in order]
        flat_idx, seg_start, seg_len, cls_rows, buckets, pos = [], [], [], [], [], 0
        for i in order:
            L_, _, _, idx, _ = seqs[i]
            if L_ == 0:
                seg_start.append(pos); seg_len.append(0); cls_rows.append(0)
                continue
            if buckets and buckets[-1][1] == L_:
                buckets[-1][0] += 1
            else:
                buckets.append([1, L_])
            flat_idx += idx
            seg_start.append(pos + n_glob); seg_len.append(L_ - n_glob); cls_rows.append(pos)
            pos += L_
        dev = tokens.device
        x = ops.gather_rows(tokens.reshape(B * N, D), torch.tensor(flat_idx, dtype=torch.int32, device=dev))
        if x.shape[0] > 0:
            self.dino.block_rows(x, [b[0] for b in buckets], [b[1] for b in buckets], layer=-1)
        if return_type == "cls":
            res = ops.gather_rows(x, torch.tensor(cls_rows, dtype=torch.int32, device=dev))
        else:
            res = ops.segment_mean(x, torch.tensor(seg_start, dtype=torch.int32, device=dev), torch.tensor(seg_len, dtype=torch.int32, device=dev))
        inv = torch.empty(len(order), dtype=torch.long)
        inv[torch.tensor(order)] = torch.arange(len(order))
        return res[inv.to(dev)].reshape(B, R, D)

    def caption_bboxes(self, imgs, bboxes, capt_type: str = "cls_capt", crop_boxes: bool = False, compute_scores: bool = False):
        """model.py:1356-1390 (the crop-and-recaption baseline): ``imgs`` is a list of PIL images, ``bboxes`` [B,R,4] xywh in their
        pixels; every box is cropped (PIL), sent through ``image_transforms_no_crop`` (or ``image_transforms`` with crop_boxes),
        and captioned from its CLS token ('cls_capt') or its attention-weighted patch average ('avg_self_attn_capt').
        All B*R crops go through one forward (the reference runs R chunks of B; captions do not depend on the chunking)."""
        if capt_type not in ("cls_capt", "avg_self_attn_capt"):
            raise ValueError(f"capt_type={capt_type!r}: expected 'cls_capt' or 'avg_self_attn_capt'")
        bs, n = len(imgs), bboxes.shape[1]
        tf = self.image_transforms if crop_boxes else self.image_transforms_no_crop
        crops = torch.stack([tf(img.crop((x, y, x + w, y + h))) for img, bb in zip(imgs, bboxes.tolist()) for (x, y, w, h) in bb])
        out = self.forward(crops, get_cls_capt=capt_type == "cls_capt", get_avg_self_attn_capt=capt_type == "avg_self_attn_capt",
                           compute_scores=compute_scores)
        capts = out[capt_type]
        ret = {"bbox_capts": [capts[i * n:(i + 1) * n] for i in range(bs)]}
        if compute_scores:
            sc = out[capt_type + "_scores"]
            ret["bbox_scores"] = [sc[i * n:(i + 1) * n] for i in range(bs)]
        return ret

    # ------------------------------------------------------------------------------------------ user-level surface
    def preprocess(self, images, keep_img_ratio: bool = True, on_device: bool = False) -> torch.Tensor:
        """PIL images -> normalised fp32 batch [B,3,crop,crop].  keep_img_ratio=True: resize the short side + centre crop
        (``image_transforms``, eval_densecap.py:313-316); False: squash to a square (``image_transforms_no_crop``, :339-342).
        Boxes must be adjusted accordingly by the caller (adjust_bbox_for_transform / ..._no_scale, bbox_utils.py:170-250).
        on_device=True runs the same pipeline in ``pio_preprocess`` (bit-identical output, result stays on the GPU; images may
        also be uint8 [H,W,3] arrays / tensors); the default runs torchvision on the host like the reference."""
        if on_device:
            from .preprocess import preprocess_images

            return preprocess_images(images, self.device, self.resize_dim, self.crop_dim, keep_img_ratio)
        tf = self.image_transforms if keep_img_ratio else self.image_transforms_no_crop
        return torch.stack([tf(im) for im in images])

    def adjust_bboxes(self, sizes, bboxes, keep_img_ratio: bool = True) -> torch.Tensor:
        """Boxes of the ORIGINAL images ((width, height) per image, xywh pixels) -> the coordinates of ``preprocess(...,
        keep_img_ratio)``: the vectorised adjust_bbox_for_transform / ..._no_scale of the eval drivers (bbox_utils.py:170-250)."""
        from .boxes import adjust_bboxes

        return adjust_bboxes(sizes, bboxes, self.resize_dim, self.crop_dim, keep_img_ratio)

    def caption(self, imgs, caption_from: str = "patches", bboxes=None, traces=None, masks=None, region_sets: bool = False,
                use_gaussian_weighting: bool = False, gaussian_variance: float = 1.0, use_attention_weighting: bool = False,
                compute_scores: bool = False, return_ids: bool = False):
        """The eval drivers' / HF wrapper's vocabulary mapped onto ``forward``:

        caption_from='cls' -> get_cls_capt; 'avg_self_attn' -> get_avg_self_attn_capt; 'patches' -> the region inputs decide:
        ``bboxes`` (one caption per box, or per box SET with region_sets=True), ``traces``, ``masks``, or -- with no region at
        all -- the whole-image patch average (get_avg_patch_capt).  use_gaussian_weighting -> gaussian_avg +
        gaussian_bbox_variance (boxes) / gaussian_img_variance (whole image); use_attention_weighting ->
        use_attn_map_for_bboxes (boxes) / use_attention_tracing (traces)."""
        if caption_from not in ("patches", "cls", "avg_self_attn"):
            raise ValueError(f"caption_from={caption_from!r}: expected 'patches', 'cls' or 'avg_self_attn'")
        kw = dict(get_cls_capt=caption_from == "cls", get_avg_self_attn_capt=caption_from == "avg_self_attn",
                  compute_scores=compute_scores, return_ids=return_ids)
        if caption_from == "patches":
            if bboxes is None and traces is None and masks is None:
                kw.update(get_avg_patch_capt=True, gaussian_img_variance=gaussian_variance if use_gaussian_weighting else 100)
            else:
                kw.update(bboxes=bboxes, traces=traces, masks=masks, get_controllable_capts=region_sets,
                          gaussian_avg=use_gaussian_weighting, gaussian_bbox_variance=gaussian_variance,
                          use_attn_map_for_bboxes=use_attention_weighting and bboxes is not None,
                          use_attention_tracing=use_attention_weighting and traces is not None)
        return self.forward(imgs, **kw)

    def forward_pipelined(self, batches, overlap_compute: bool = True, **flags):
        """Serving loop over host-resident batches: yields ``forward(**batch, **flags)`` for every batch, in order.

        ``batches`` is an iterable of dicts (``imgs`` plus optional ``bboxes`` / ``masks`` tensors, ``traces`` lists), ideally in
        pinned memory.  Two things run under the kernels of batch i: the host->device copy of batch i+1 (copy stream, two
        persistent staging slots) and -- with ``overlap_compute`` -- the forward of batch i+1 itself, issued on a second
        compute stream with its own scratch buffers, so that its large ViT kernels fill the SMs the small decode kernels of
        batch i leave idle.  Same outputs as calling ``forward`` batch by batch."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            # high priority: the copies get their own hardware queue and are never stuck behind a forward's ~1300 queued kernels
            self._copy_stream = torch.cuda.Stream(self.device, priority=-1)
            self._alt_stream = torch.cuda.Stream(self.device)
            self._stage_bufs, self._stage_free = {}, {}
        copy = self._copy_stream
        streams = (main, self._alt_stream) if overlap_compute else (main, main)
        if overlap_compute:
            # two forwards in flight: a small-batch decode (one persistent cooperative kernel) takes half of the SMs, so that the
            # other batch's kernels -- or its decode -- run beside it instead of queueing behind a grid that owns every SM
            L.lib().pio_set_decode_fused(-1, 74)

        def stage(batch, slot):
            """copy `batch` into the persistent device buffers of `slot` (allocated once per shape: no allocator traffic,
            no implicit synchronisation in the loop)"""
            dev_batch, ev = {}, torch.cuda.Event()
            with torch.cuda.stream(copy):
                if slot in self._stage_free:
                    copy.wait_event(self._stage_free[slot])  # the forward that last read this slot has finished
                for k, v in batch.items():
                    if torch.is_tensor(v):
                        key = (slot, k, tuple(v.shape), v.dtype)
                        buf = self._stage_bufs.get(key)
                        if buf is None:
                            buf = self._stage_bufs[key] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                        buf.copy_(v, non_blocking=True)
                        dev_batch[k] = buf
                    else:
                        dev_batch[k] = v
                ev.record(copy)
            return dev_batch, ev

        # Strings without a stall: a plain forward() ends in ids.cpu() + detokenisation, which would hold the host until batch i
        # is finished and leave the GPU idle while batch i+1 is being launched.  Here the forward runs with return_ids=True,
        # the id matrices are copied to pinned host buffers on the batch's own stream, and the (batched) detokenisation of
        # batch i happens at consumption time -- under the kernels of batch i+1.
        defer_text = (not flags.get("return_ids", False) and self.viecap is None and not self.calculate_argmax_text
                      and flags.get("caption_bboxes_type") is None)
        run_flags = dict(flags, return_ids=True) if defer_text else flags
        if getattr(self, "_ids_host", None) is None:
            self._ids_host = {}

        def launch(i, staged):
            """issue the forward of batch i on its compute stream; returns (outputs, completion event, pinned id copies)"""
            cur, ev = staged
            st = streams[i & 1]
            host_ids = {}
            with torch.cuda.stream(st):
                st.wait_event(ev)
                if i == 1 and first_done:
                    st.wait_event(first_done[0])  # lazily built device caches (pos-embed ...) of the very first forward are complete
                out = self.forward(**cur, **run_flags)
                if defer_text:
                    for k, v in out.items():
                        if torch.is_tensor(v) and v.dtype == torch.int32:
                            key = (i & 1, k, tuple(v.shape))
                            hb = self._ids_host.get(key)
                            if hb is None:
                                hb = self._ids_host[key] = torch.empty(v.shape, dtype=torch.int32, pin_memory=True)
                            hb.copy_(v, non_blocking=True)
                            host_ids[k] = hb
                done = torch.cuda.Event()
                done.record(st)
            self._stage_free[i & 1] = done
            if i == 0:
                first_done.append(done)
            return out, done, host_ids

        def texts(out, host_ids):
            """the ids of a finished batch -> the reference's string outputs ([B][R] lists / [B] lists)"""
            res = dict(out)
            for k, hb in host_ids.items():
                flat = self._host_ids_to_text(hb.reshape(-1, hb.shape[-1])) if hb.numel() else []
                if hb.dim() == 3:
                    g = hb.shape[1]
                    res[k] = [flat[b * g:(b + 1) * g] for b in range(hb.shape[0])]
                else:
                    res[k] = flat
            return res

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        i = 0
        first_done: list = []
        try:
            yield from self._pipeline_loop(it, nxt, launch, stage, main, defer_text, texts, first_done)
        finally:
            if overlap_compute:
                L.lib().pio_set_decode_fused(-1, 0)

    def _pipeline_loop(self, it, nxt, launch, stage, main, defer_text, texts, first_done):
        i = 0
        inflight = launch(0, stage(nxt, 0))
        while inflight is not None:
            nxt = next(it, None)
            following = launch(i + 1, stage(nxt, (i + 1) & 1)) if nxt is not None else None  # issued before batch i is consumed
            out, done, host_ids = inflight
            main.wait_event(done)  # the caller reads the results on the current stream
            for v in out.values():
                if torch.is_tensor(v):
                    v.record_stream(main)
            if defer_text:
                done.synchronize()  # the pinned id copies of batch i are complete (batch i+1 is already queued)
                out = texts(out, host_ids)
            inflight = following
            i += 1
            yield out
