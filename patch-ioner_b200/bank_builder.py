"""Caption-memory builder, device side: everything of `Im2TxtProjector._build_memory`
(Patch-ioner/src/decap/im2txtprojection/im2txtprojection.py:448-560) that comes AFTER the CLIP text encoder --

  * the Talk2DINO projection of the CLIP text features (`ProjectionLayer.project_clip_txt`, src/talk2dino/talk2dino.py:73-83:
    Linear -> [act -> Linear]*), as fp32 GEMMs of libpio_sm100 with the activation in the epilogue (:519-523);
  * the reference's file name scheme (:234) and on-disk layout: HDF5 datasets '{name}-embeddings' fp32 [M,D] and '{name}-text'
    utf-8 (:543-555) when h5py is importable, else the same two arrays in a torch file that `Patchioner(memory_bank=...)`
    reads as well;
  * row shards for a bank spread over ranks (`shard_rows`; the projection of a sharded bank is `patchioner_b200.dist`).

The CLIP text encoder itself (weights + tokenizer from the network) is outside the hot path: the builder takes its output
features.  There is no CPU fallback for the MLP."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops

_ACTS = {None: L.ACT_NONE, "tanh": L.ACT_TANH, "relu": L.ACT_RELU}


def talk2dino_project(clip_text_features: torch.Tensor, state_dict: Dict[str, torch.Tensor], act: Optional[str] = "tanh",
                      device="cuda", rows_per_call: int = 262144) -> torch.Tensor:
    """`project_clip_txt` for [M, clip_dim] text features -> fp32 [M, dino_dim] on the host.  `state_dict` uses the reference's
    keys (`linear_layer.*`, `hidden_layers.{i}.*`; the old `linear_layer2.*` alias is accepted, talk2dino.py:85-91)."""
    if act not in _ACTS:
        raise ValueError(f"unknown activation {act!r} (talk2dino.py:44-52 knows tanh / relu / sigmoid; sigmoid is not built)")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise L.PioError("the bank builder runs its MLP on CUDA only: there is no CPU fallback")
    sd = dict(state_dict)
    if "linear_layer2.weight" in sd:
        sd["hidden_layers.0.weight"], sd["hidden_layers.0.bias"] = sd.pop("linear_layer2.weight"), sd.pop("linear_layer2.bias")
    w = {k: v.detach().to(dev, torch.float32).contiguous() for k, v in sd.items() if k.startswith(("linear_layer.", "hidden_layers."))}
    n_hidden = len([k for k in w if k.startswith("hidden_layers.") and k.endswith(".weight")])
    out = []
    feats = clip_text_features
    for s in range(0, feats.shape[0], rows_per_call):
        x = feats[s:s + rows_per_call].to(dev, torch.float32).contiguous()  # `.float()` of :74
        # the activation precedes every hidden layer (:78-81): it goes into the epilogue of the layer before
        x = ops.linear(x, w["linear_layer.weight"], "fp32", bias=w["linear_layer.bias"],
                       act=_ACTS[act] if n_hidden > 0 else L.ACT_NONE)
        for i in range(n_hidden):
            last = i == n_hidden - 1
            x = ops.linear(x, w[f"hidden_layers.{i}.weight"], "fp32", bias=w[f"hidden_layers.{i}.bias"],
                           act=L.ACT_NONE if last else _ACTS[act])
        out.append(x.cpu())
    return torch.cat(out, 0) if out else torch.empty(0, w["linear_layer.weight"].shape[0])


def bank_filename(dataset_name: str, clip_modelname: str, support_memory_size: int, prefix: str = "", postfix: str = "",
                  talk2dino_attn_type_str: str = "") -> str:
    """im2txtprojection.py:234."""
    return (prefix + f"{dataset_name}_text_embeddings{talk2dino_attn_type_str}{postfix}-{clip_modelname.replace('/', '.')}"
            f"-{support_memory_size}.h5")


def write_bank(path: str, embeddings: torch.Tensor, texts: Sequence[str], name: str = "coco") -> str:
    """Store the bank in the reference's layout.  Returns the path written (``.h5`` needs h5py; without it the same two
    arrays go to ``<path>.pt``, which `Patchioner(memory_bank=...)` accepts as well)."""
    emb = embeddings.detach().to("cpu", torch.float32).contiguous()
    if len(texts) != emb.shape[0]:
        raise ValueError(f"{emb.shape[0]} embeddings but {len(texts)} captions")  # the assert of :536
    try:
        import h5py
    except ImportError:
        h5py = None
    if path.endswith((".h5", ".hdf5")) and h5py is not None:
        with h5py.File(path, "w") as hf:
            hf.create_dataset(f"{name}-embeddings", data=emb.numpy(), dtype="float32")
            ds = hf.create_dataset(f"{name}-text", shape=(len(texts),), dtype=h5py.string_dtype(encoding="utf-8"))
            for i, t in enumerate(texts):
                ds[i] = t
        return path
    if path.endswith((".h5", ".hdf5")):
        path = path + ".pt"
    torch.save({f"{name}-embeddings": emb, f"{name}-text": list(texts)}, path)
    return path


def read_bank(path: str) -> Tuple[torch.Tensor, Optional[List[str]]]:
    """(embeddings fp32 [M,D], captions) from a file written by `write_bank` (torch flavour) or by the reference (HDF5)."""
    if path.endswith(".pt"):
        d = torch.load(path, map_location="cpu", weights_only=False)
        if torch.is_tensor(d):
            return d.float(), None
        ek = [k for k in d if k.endswith("-embeddings")][0]
        tk = [k for k in d if k.endswith("-text")]
        return d[ek].float(), (list(d[tk[0]]) if tk else None)
    import h5py
    with h5py.File(path, "r") as hf:
        ek = [k for k in hf.keys() if k.endswith("-embeddings")][0]
        tk = [k for k in hf.keys() if k.endswith("-text")]
        texts = [t.decode() if isinstance(t, bytes) else t for t in hf[tk[0]][:]] if tk else None
        return torch.from_numpy(hf[ek][:]).float(), texts


def shard_rows(M: int, world: int, rank: int) -> Tuple[int, int]:
    """[start, end) of the bank rows rank `rank` keeps when the bank is row-sharded over `world` ranks (BASELINE configs[4])."""
    per = (M + world - 1) // world
    return min(rank * per, M), min((rank + 1) * per, M)


def build_bank(clip_text_features: torch.Tensor, texts: Sequence[str], talk2dino_state_dict: Optional[Dict[str, torch.Tensor]] = None,
               act: Optional[str] = "tanh", out_path: Optional[str] = None, name: str = "coco", device="cuda"):
    """CLIP text features (+ Talk2DINO weights) -> bank rows; optionally written to disk.  Zero rows are kept, exactly as the
    reference stores them (they are filtered when the bank is loaded, im2txtprojection.py:342-345)."""
    emb = clip_text_features.float().cpu() if talk2dino_state_dict is None else \
        talk2dino_project(clip_text_features, talk2dino_state_dict, act, device)
    written = write_bank(out_path, emb, texts, name) if out_path else None
    return emb, written
