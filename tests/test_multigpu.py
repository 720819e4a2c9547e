"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): the reference's own way of using several GPUs,
`torch.nn.DataParallel(model)` (main_spectrogram.py:118-121: one Python thread per GPU calls forward on a
replica), and the one-process-per-GPU sharding of SURVEY 8e over NCCL.  Sequences are independent, so every
result must be bit-identical to the single-GPU one."""
import os
import socket

import pytest
import torch

from tests import fixtures as fx

pytestmark = pytest.mark.gpu


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_dataparallel_replicas_match_single_gpu():
    _need_two()
    from skeleton_action_recognition_b200 import VirtualRadar
    from skeleton_action_recognition_b200.models.resnet import Model
    layer = VirtualRadar(wavelength=5e-4, device="cuda:0").to("cuda:0")
    x = fx.s1_iid(37).cuda(0)                        # odd batch: shards of 19 and 18
    want = layer(x)
    dp = torch.nn.DataParallel(layer, device_ids=[0, 1])
    for _ in range(3):                               # threads launch concurrently, one per device
        got = dp(x)
        assert got.device == want.device and torch.equal(got, want)
    model = Model(num_filters=8, image_size=64, base_model=torch.nn.Flatten()).to("cuda:0")
    dpm = torch.nn.DataParallel(model, device_ids=[0, 1])
    assert torch.equal(dpm(x), model(x))
    # trainable radar parameters: DataParallel reduces the replicas' gradients onto device 0
    tl = VirtualRadar(wavelength=5e-3, train_wavelength=True, train_radar_location=True, device="cuda:0").to("cuda:0")
    xs = fx.s3_smooth(8, T=300).cuda(0)
    tl(xs).square().mean().backward()
    g1 = (tl.wavelength.grad.clone(), tl.radar_location.grad.clone())
    tl.zero_grad()
    torch.nn.DataParallel(tl, device_ids=[0, 1])(xs).square().mean().backward()
    assert torch.allclose(tl.wavelength.grad, g1[0], rtol=1e-4) and torch.allclose(tl.radar_location.grad, g1[1], rtol=1e-4, atol=1e-6 * g1[1].abs().max())


def test_trained_stft_kernels_on_the_second_device_and_under_dataparallel():
    """The tcgen05 STFT path (per-device kernel attributes, tensor-memory allocation, bulk copies) on cuda:1, and the
    DataParallel replicas of a layer with trainable kernels: forward bits and kernel gradients as on one GPU."""
    _need_two()
    from skeleton_action_recognition_b200 import VirtualRadar
    x = fx.s1_iid(12)
    l0 = VirtualRadar(wavelength=5e-4, train_stft_kernel=True, device="cuda:0").to("cuda:0")
    l1 = VirtualRadar(wavelength=5e-4, train_stft_kernel=True, device="cuda:1").to("cuda:1")
    y0 = l0(x.cuda(0))
    y1 = l1(x.cuda(1))
    assert y1.device.index == 1 and torch.equal(y0.cpu(), y1.cpu())
    y0.square().mean().backward()
    y1.square().mean().backward()
    assert torch.equal(l0.stft.wsin.grad.cpu(), l1.stft.wsin.grad.cpu())
    g_single = l0.stft.wcos.grad.clone()
    l0.zero_grad()
    dp = torch.nn.DataParallel(l0, device_ids=[0, 1])
    yd = dp(x.cuda(0))
    assert torch.equal(yd, y0.detach())
    yd.square().mean().backward()                    # two shards' kernel gradients, reduced onto device 0
    assert torch.allclose(l0.stft.wcos.grad, g_single, rtol=1e-3, atol=1e-5 * float(g_single.abs().max()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from skeleton_action_recognition_b200 import VirtualRadar, sharded_forward
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        layer = VirtualRadar(wavelength=5e-4, device="cuda:%d" % rank).to("cuda:%d" % rank)
        x = fx.s1_iid(75).cuda(rank)                 # 38 + 37
        full = sharded_forward(layer, x)
        torch.save(full.cpu(), os.path.join(out_dir, "rank%d.pt" % rank))
        if rank == 0:
            torch.save(layer(x).cpu(), os.path.join(out_dir, "single.pt"))
    finally:
        dist.destroy_process_group()


def test_one_process_per_gpu_sharding_over_nccl(tmp_path):
    _need_two()
    import torch.multiprocessing as mp
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    single = torch.load(tmp_path / "single.pt")
    for r in range(2):
        assert torch.equal(torch.load(tmp_path / ("rank%d.pt" % r)), single)
