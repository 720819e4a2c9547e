"""Input side and the consumer module (8f-1): the `utils.Dataset` mirror returns raw samples
and labels from the reference's file formats (CPU tests); on the GPU the up-sampled batches equal what the
reference's `Dataset.__getitem__` produces (oracle.pad_frames.dataset_getitem = scipy, as in utils.py:128-140),
and `models.resnet.Model` is VirtualRadar -> unsqueeze -> nearest resize -> classifier (models/resnet.py:23-28)."""
import os
import pickle

import numpy as np
import pytest
import torch


def _write_dataset(tmp_path, n=6, T=40, V=25, M=2, seed=0):
    rng = np.random.default_rng(seed)
    data = (rng.standard_normal((n, 3, T, V, M)) * 0.3).astype(np.float32)
    labels = [int(v) for v in rng.integers(0, 60, n)]
    names = ["S%03d" % i for i in range(n)]
    np.save(tmp_path / "train_data_joint.npy", data)           # data_gen/gen_joint_data.py:138-151 layout
    with open(tmp_path / "train_label.pkl", "wb") as f:
        pickle.dump((names, labels), f)
    return data, labels


def test_dataset_reads_reference_file_formats(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    from skeleton_action_recognition_b200.feeder import Dataset
    data, labels = _write_dataset(tmp_path)
    ds = Dataset(tmp_path / "train_data_joint.npy", tmp_path / "train_label.pkl", num_pad_frames=7, sigma=2)
    assert len(ds) == 6 and ds.T == 40 and ds.num_pad_frames == 7 and ds.sigma == 2
    x, y = ds[3]
    assert x.dtype == torch.float32 and tuple(x.shape) == (3, 40, 25, 2) and int(y) == labels[3]
    assert np.array_equal(x.numpy(), data[3])
    loader = torch.utils.data.DataLoader(ds, batch_size=4, shuffle=False)
    xb, yb = next(iter(loader))
    assert tuple(xb.shape) == (4, 3, 40, 25, 2) and yb.tolist() == labels[:4]
    with pytest.raises(FileNotFoundError):
        Dataset(tmp_path / "missing.npy", tmp_path / "train_label.pkl")
    with pytest.raises(RuntimeError):
        ds.upsample(xb)                                           # no CPU path


@pytest.mark.gpu
def test_gpu_batches_equal_reference_getitem(tmp_path):
    from oracle import pad_frames as opf
    from skeleton_action_recognition_b200.feeder import Dataset, gpu_batches
    data, labels = _write_dataset(tmp_path, n=5, T=48)
    ds = Dataset(tmp_path / "train_data_joint.npy", tmp_path / "train_label.pkl", num_pad_frames=9, sigma=3)
    loader = torch.utils.data.DataLoader(ds, batch_size=3, shuffle=False, pin_memory=True)
    seen = 0
    for xb, yb in gpu_batches(loader, "cuda:0"):
        assert xb.is_cuda and tuple(xb.shape[1:]) == (3, 9 * 48, 25, 2)
        for i in range(xb.shape[0]):
            want = opf.dataset_getitem(data[seen], 9, 3).numpy()
            got = xb[i].cpu().numpy()
            assert np.array_equal(got, want)                       # bit-equal to the reference's __getitem__
            assert int(yb[i]) == labels[seen]
            seen += 1
    assert seen == 5
    raw = next(iter(gpu_batches(loader, "cuda:0", upsample=False)))[0]
    assert tuple(raw.shape) == (3, 3, 48, 25, 2)
    # double buffering (the default) hands out the same batches as the plain loop, in order
    a = [(x.cpu(), y.cpu()) for x, y in gpu_batches(loader, "cuda:0", num_pad_frames=5)]
    b = [(x.cpu(), y.cpu()) for x, y in gpu_batches(loader, "cuda:0", num_pad_frames=5, prefetch=False)]
    assert len(a) == len(b) == 2 and all(torch.equal(p[0], q[0]) and torch.equal(p[1], q[1]) for p, q in zip(a, b))


def test_model_state_dict_matches_the_reference_model():
    """Keys and shapes of `Model.state_dict()` are those of the reference's `models/resnet.py` Model: the classifier
    (models/resnet18.py, recorded by loading that file by path in the build container; tests/golden/
    resnet18_state_dict_f16.json for num_filters=16) under `base_model.`, the radar layer's four entries under
    `virtual_radar.` -- so checkpoints of the reference load here and the other way round."""
    import json
    from skeleton_action_recognition_b200.models.resnet import Model
    want = {"base_model." + k: tuple(v) for k, v in
            json.load(open(os.path.join(os.path.dirname(__file__), "golden", "resnet18_state_dict_f16.json"))).items()}
    want.update({"virtual_radar.wavelength": (), "virtual_radar.radar_location": (3,),
                 "virtual_radar.stft.wsin": (256, 1, 256), "virtual_radar.stft.wcos": (256, 1, 256)})
    model = Model(num_classes=60, num_filters=16, device="cpu")
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    twin = Model(num_classes=60, num_filters=16, device="cpu")
    twin.load_state_dict(model.state_dict())                      # strict


@pytest.mark.gpu
def test_model_is_radar_resize_classifier():
    from skeleton_action_recognition_b200.models.resnet import Model
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(3, 3, 300, 25, 2, generator=g) * 0.3).cuda()
    model = Model(num_classes=60, num_filters=8, image_size=128).cuda().eval()
    with torch.no_grad():
        logits = model(x)
        img = torch.nn.functional.interpolate(model.virtual_radar(x).unsqueeze(1), 128)   # the reference's three steps
        assert torch.equal(model.spectrogram_image(x), img)
        assert torch.equal(logits, model.base_model(img))
    assert tuple(logits.shape) == (3, 60)
    # one training step of the consumer (config 5): gradients reach the classifier, the radar is forward-only
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss = torch.nn.functional.cross_entropy(model(x), torch.tensor([1, 2, 3], device="cuda"))
    loss.backward()
    opt.step()
    assert torch.isfinite(loss) and model.base_model.conv1.weight.grad is not None
    assert sorted(k for k in model.state_dict() if k.startswith("virtual_radar")) == [
        "virtual_radar.radar_location", "virtual_radar.stft.wcos", "virtual_radar.stft.wsin", "virtual_radar.wavelength"]


@pytest.mark.gpu
def test_model_upsamples_raw_sequences_on_the_device():
    from skeleton_action_recognition_b200 import pad_frames
    from skeleton_action_recognition_b200.models.resnet import Model
    g = torch.Generator().manual_seed(6)
    x = (torch.randn(2, 3, 60, 25, 2, generator=g) * 0.3).cuda()
    model = Model(num_filters=8, image_size=64, num_pad_frames=20, base_model=torch.nn.Flatten()).cuda()
    img = model.spectrogram_image(x)
    want = torch.nn.functional.interpolate(model.virtual_radar(pad_frames(x, 20, 3)).unsqueeze(1), 64)
    assert torch.equal(img, want) and tuple(model(x).shape) == (2, 64 * 64)
