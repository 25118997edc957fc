"""Pre-encoded latent pipeline: the data format either side of the frozen-encoder step of LDM training.

Writer  = `pre_encode.py:39-125` (`PreEncodedLatentsInferenceWrapper.validation_step`): one `{rank:03d}{batch:06d}{i:04d}.npy` ([C, N]
          fp32 latents) + `.json` (metadata with the padding mask nearest-interpolated to the latent length) per clip under
          `<output>/<rank>/`, plus `<output>/details.json`.
Reader  = `data/dataset.py:265-360` (`PreEncodedDataset`): crop to `latent_crop_length` (random start inside the un-padded part when
          `random_crop`), min/max length filtering, `info["audio"] = latents`.
The on-disk format is the reference's, byte-compatible in both directions, so a dataset pre-encoded by either side trains the other
(`pre_encoded: true`, training/diffusion.py:344, 376-379).  Encoding runs on `OobleckEngine.encode_audio` (8 ms per 47 s clip on B200).
"""
import json
import os
import random

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


def _latent_files(path, ext):
    out = []
    for root, _, files in os.walk(path):
        for fn in files:
            if fn.endswith("." + ext) and not fn.startswith("."):
                out.append(os.path.join(root, fn))
    return sorted(out)


class PreEncodedDataset(torch.utils.data.Dataset):
    """`stable_audio_tools.data.dataset.PreEncodedDataset` semantics over plain directory paths."""

    def __init__(self, paths, latent_crop_length=None, min_length_sec=None, max_length_sec=None, random_crop=False, latent_extension="npy",
                 custom_metadata_fns=None):
        super().__init__()
        if isinstance(paths, (str, os.PathLike)):
            paths = [paths]
        self.latent_extension = latent_extension
        self.filenames = []
        for p in paths:
            self.filenames.extend(_latent_files(str(p), latent_extension))
        self.custom_metadata_fns = dict(custom_metadata_fns or {})
        self.latent_crop_length = latent_crop_length
        self.random_crop = random_crop
        self.min_length_sec, self.max_length_sec = min_length_sec, max_length_sec

    def __len__(self):
        return len(self.filenames)

    def __getitem__(self, idx):
        fn = self.filenames[idx]
        latents = torch.from_numpy(np.load(fn))   # [C, N]
        with open(fn[: -len(self.latent_extension) - 1] + ".json") as f:
            info = json.load(f)
        info["latent_filename"] = fn
        if self.latent_crop_length is not None:
            pm = info["padding_mask"]
            last_ix = len(pm) - 1 - pm[::-1].index(1)
            start = random.randint(0, last_ix - self.latent_crop_length) if (self.random_crop and last_ix > self.latent_crop_length) else 0
            latents = latents[:, start:start + self.latent_crop_length]
            info["padding_mask"] = pm[start:start + self.latent_crop_length]
            info["latent_crop_length"] = self.latent_crop_length
            info["latent_crop_start"] = start
        info["padding_mask"] = [torch.tensor(info["padding_mask"])]
        seconds_total = info.get("seconds_total")
        if seconds_total is not None:
            if self.min_length_sec is not None and seconds_total < self.min_length_sec:
                return self[random.randrange(len(self))]
            if self.max_length_sec is not None and seconds_total > self.max_length_sec:
                return self[random.randrange(len(self))]
        for root, fn_ in self.custom_metadata_fns.items():
            if root in fn:
                info.update(fn_(info, None))
            if info.get("__reject__"):
                return self[random.randrange(len(self))]
            if info.get("__replace__") is not None:
                latents = info["__replace__"]
        info["audio"] = latents
        return latents, info
