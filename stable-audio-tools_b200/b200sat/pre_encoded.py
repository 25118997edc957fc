"""Pre-encoded latent pipeline: the data format either side of the frozen-encoder step of LDM training.

Writer  = `pre_encode.py:39-125` (`PreEncodedLatentsInferenceWrapper.validation_step`): one `{rank:03d}{batch:06d}{i:04d}.npy` ([C, N]
          fp32 latents) + `.json` (metadata with the padding mask nearest-interpolated to the latent length) per clip under
          `<output>/<rank>/`, plus `<output>/details.json`.
Reader  = the reference's own `PreEncodedDataset` (`data/dataset.py:265-360`), imported unchanged (`reference_dataset()` below).
The on-disk format is the reference's, so a dataset pre-encoded here trains the unmodified reference and vice versa
(`pre_encoded: true`, training/diffusion.py:344, 376-379).  Encoding runs on `OobleckEngine.encode_audio` (8 ms per 47 s clip on B200).
"""
import json
import os

import numpy as np
import torch


def _nearest_resize_mask(mask, size):
    """F.interpolate(mask[None, None].float(), size=size, mode='nearest').int() without the round trip: src = floor(i * in / out)."""
    mask = torch.as_tensor(mask)
    n = mask.shape[-1]
    idx = torch.div(torch.arange(size, dtype=torch.float32) * (float(n) / float(size)), 1, rounding_mode="floor").long().clamp_(max=n - 1)
    return mask[..., idx].to(torch.int32)


def write_details(output_path, model_config=None, dataset_config=None, sample_size=None, args=None):
    os.makedirs(output_path, exist_ok=True)
    p = os.path.join(output_path, "details.json")
    if not os.path.exists(p):
        with open(p, "w") as f:
            json.dump({"model_config": model_config, "dataset_config": dataset_config, "sample_size": sample_size, "args": args}, f)


def write_pre_encoded(encode_fn, audio, metadata, output_path, rank=0, batch_idx=0):
    """Encode one batch and write it in the reference layout.  encode_fn: audio [B, C, T] -> latents [B, L, N] (e.g.
    `lambda a: engine.encode_audio(a) ` scaled like the pretransform); metadata: list of dicts with at least `padding_mask` [T]."""
    if audio.ndim == 4 and audio.shape[0] == 1:
        audio = audio[0]
    with torch.no_grad():
        latents = encode_fn(audio)
    latents = latents.detach().float().cpu().numpy()
    d = os.path.join(output_path, str(rank))
    os.makedirs(d, exist_ok=True)
    paths = []
    for i, latent in enumerate(latents):
        latent_id = f"{rank:03d}{batch_idx:06d}{i:04d}"
        with open(os.path.join(d, latent_id + ".npy"), "wb") as f:
            np.save(f, latent)
        md = dict(metadata[i])
        md["padding_mask"] = _nearest_resize_mask(md["padding_mask"], latent.shape[1]).cpu().numpy().tolist()
        for k, v in list(md.items()):
            if isinstance(v, torch.Tensor):
                md[k] = v.cpu().numpy().tolist()
        with open(os.path.join(d, latent_id + ".json"), "w") as f:
            json.dump(md, f)
        paths.append(os.path.join(d, latent_id + ".npy"))
    return paths


def reference_dataset(paths, **kw):
    """The READER is the reference's own class (`stable_audio_tools.data.dataset.PreEncodedDataset`, data/dataset.py:265-360) — the data
    layer is out of the hot path's scope and is reused as-is; this helper only builds its `LocalDatasetConfig` list from plain paths.
    Needs the reference package importable (and its `webdataset` dependency, or a stand-in module of that name)."""
    from stable_audio_tools.data.dataset import LocalDatasetConfig, PreEncodedDataset
    if isinstance(paths, (str, os.PathLike)):
        paths = [paths]
    cfgs = [LocalDatasetConfig(id=f"pre_encoded_{i}", path=str(p)) for i, p in enumerate(paths)]
    return PreEncodedDataset(cfgs, **kw)
