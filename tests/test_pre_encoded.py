"""Pre-encoded latent files: writer/reader round trip in the reference's on-disk format (pre_encode.py:39-125, dataset.py:265-360)."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "stable-audio-tools_b200"))


def _load():
    import importlib.util
    p = os.path.join(os.path.dirname(__file__), "..", "stable-audio-tools_b200", "b200sat", "pre_encoded.py")
    spec = importlib.util.spec_from_file_location("b200sat_pre_encoded", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_padding_mask_resize_matches_interpolate_nearest():
    pe = _load()
    g = torch.Generator().manual_seed(0)
    for n, size in ((2097152, 1024), (65536, 32), (1000, 37), (37, 1000), (5, 5)):
        mask = (torch.rand(n, generator=g) > 0.3).int()
        ref = F.interpolate(mask[None, None].float(), size=size, mode="nearest").squeeze().int()
        assert torch.equal(pe._nearest_resize_mask(mask, size), ref), (n, size)


def test_write_then_read_reference_layout(tmp_path):
    pe = _load()
    g = torch.Generator().manual_seed(1)
    audio = torch.randn(3, 2, 4096, generator=g)
    enc = lambda a: a.reshape(3, 2, 64, 64).mean(-1).repeat(1, 32, 1)   # stand-in encoder: [B, 64, 64]
    md = [{"padding_mask": torch.cat([torch.ones(4096 - 512 * i), torch.zeros(512 * i)]).int(), "seconds_total": 10.0 + i, "seconds_start": 0,
           "prompt": f"clip {i}"} for i in range(3)]
    pe.write_details(str(tmp_path), model_config={"a": 1}, sample_size=4096)
    paths = pe.write_pre_encoded(enc, audio, md, str(tmp_path), rank=2, batch_idx=7)
    assert [os.path.basename(p) for p in paths] == ["0020000070000.npy", "0020000070001.npy", "0020000070002.npy"]
    assert os.path.exists(tmp_path / "details.json") and os.path.isdir(tmp_path / "2")
    assert np.load(paths[1]).shape == (64, 64) and np.load(paths[1]).dtype == np.float32
    j = json.load(open(paths[2][:-4] + ".json"))
    assert len(j["padding_mask"]) == 64 and sum(j["padding_mask"]) == 64 - 16 and j["prompt"] == "clip 2"
    # the reader is the REFERENCE's PreEncodedDataset, unmodified (baseline/_ref): our files must satisfy it
    from baseline import ref_loader
    if not ref_loader.available():
        return
    ref_loader.load(force_sdpa=True)
    ds = pe.reference_dataset(str(tmp_path), latent_crop_length=32, random_crop=True)
    assert len(ds) == 3
    lat, info = ds[1]
    assert lat.shape == (64, 32) and info["audio"] is lat and info["padding_mask"][0].shape == (32,)
    k = paths.index(info["latent_filename"])                      # the reference does not sort its file list
    assert info["seconds_total"] == 10.0 + k and 0 <= info["latent_crop_start"] <= 64
    full = torch.from_numpy(np.load(paths[k]))
    assert torch.equal(lat, full[:, info["latent_crop_start"]:info["latent_crop_start"] + 32])
    ds2 = pe.reference_dataset([str(tmp_path)], min_length_sec=11.5)
    for i in range(3):
        assert ds2[i][1]["seconds_total"] >= 11.5
