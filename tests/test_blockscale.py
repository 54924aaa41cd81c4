"""Block-scaled e4m3 codec (utils/blockscale.py: UE8M0 scale per 32 values) and DeepFM training on the values such a
table would deliver (`table_precision="fp8_block"`)."""
import torch

from harness import run_cluster


def test_block_e4m3_roundtrip_error_and_format():
    from adapm_b200.utils.blockscale import (BLOCK, bytes_per_value, dequantize_block_e4m3, fake_quantize_block_e4m3,
                                            quantize_block_e4m3)

    g = torch.Generator().manual_seed(0)
    x = torch.randn(500, 80, generator=g) * torch.logspace(-6, 3, 500).unsqueeze(1)     # rows of very different magnitude
    q, e = quantize_block_e4m3(x)
    assert q.dtype == torch.float8_e4m3fn and e.dtype == torch.int8
    assert q.shape == (500, 3, BLOCK) and e.shape == (500, 3)                            # 80 values -> 3 blocks (padded)
    assert abs(bytes_per_value() - 33 / 32) < 1e-12
    y = dequantize_block_e4m3(q, e, 80)
    assert y.shape == x.shape
    # e4m3 has 3 mantissa bits: an element is off by at most 2^-4 of its block's power-of-two ceiling (<= 2 x block max)
    xb = torch.nn.functional.pad(x, (0, 16)).view(500, 3, BLOCK)
    yb = torch.nn.functional.pad(y, (0, 16)).view(500, 3, BLOCK)
    bound = xb.abs().amax(-1, keepdim=True) * 2.0 ** -3
    assert bool(((xb - yb).abs() <= bound + 1e-30).all())
    # the scale is a power of two and the scaled block fits e4m3
    assert bool((q.to(torch.float32).abs() <= 448).all())
    # fake quantisation is idempotent, keeps zeros, and keeps the sign
    f1 = fake_quantize_block_e4m3(x)
    assert torch.equal(fake_quantize_block_e4m3(f1), f1)
    assert float(fake_quantize_block_e4m3(torch.zeros(4, 32)).abs().max()) == 0.0
    assert bool((torch.sign(f1) * torch.sign(x) >= 0).all())


def test_fake_quant_straight_through_gradient():
    from adapm_b200.utils.blockscale import FakeQuantSTE, fake_quantize_block_e4m3

    x = torch.randn(7, 20, requires_grad=True)
    y = FakeQuantSTE.apply(x)
    assert torch.equal(y.detach(), fake_quantize_block_e4m3(x.detach()))
    (y * torch.arange(20.0)).sum().backward()
    assert torch.equal(x.grad, torch.arange(20.0).expand(7, 20))


def _ctr_worker(kv, server, wid):
    from adapm_b200.models.deepfm import DeepFM, synthetic_ctr_batch

    cfg = server._cfg
    model = DeepFM(server, kv, cfg)
    model.init_model()
    losses = []
    for s in range(60):
        ids, y = synthetic_ctr_batch(cfg, s % 6, server.my_rank())
        losses.append(model.step(ids, y))
        kv.advance_clock()
    kv.barrier()
    kv.finalize()
    return losses


def test_deepfm_trains_on_block_fp8_table_values():
    """DeepFM on the values a block-scaled e4m3 table delivers (fp32 master rows, straight-through gradients): the loss
    goes down like with the fp32 table."""
    from adapm_b200.models.deepfm import DeepFMConfig

    out = {}
    for prec in ("fp32", "fp8_block"):
        cfg = DeepFMConfig(num_features=3000, num_fields=6, embed_dim=8, hidden=(32, 32), batch_size=256, precision="fp32",
                           table_precision=prec, lr_sparse=0.05)

        def setup(server, cfg=cfg):
            server._cfg = cfg

        res = run_cluster(_ctr_worker, world=1, workers=1, mode="threads", setup_fn=setup, value_lengths=cfg.row_len,
                          num_keys=cfg.num_features)
        out[prec] = res[0][0]
    for prec, ls in out.items():
        assert sum(ls[-6:]) < 0.97 * sum(ls[:6]), (prec, ls[:6], ls[-6:])
    assert abs(sum(out["fp8_block"][-6:]) - sum(out["fp32"][-6:])) < 0.1 * sum(out["fp32"][-6:])
