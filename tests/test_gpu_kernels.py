"""GPU parity tests of the individual kernels, through the C ABI (ctypes), against the CPU oracle,
the golden fixtures minted from the live reference, and plain torch fp32 references."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import decode_oracle as D
from oracle import loss_oracle as L
from oracle import golden_inputs as GI
from tests.gpu_cases import CONV_CASES, make_conv_case, conv_reference, run_conv_case, no_tf32, r16

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ops():
    from hgb200 import ops as _ops, lib
    lib.check(lib.hg_check_device(), "hg_check_device")
    return _ops


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------ convolutions
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm(case, ops, dev):
    t = make_conv_case(case, dev)
    ref = conv_reference(t)
    out = run_conv_case(t)
    scale = float(ref.abs().max())
    # bf16 output rounding (2^-8 relative) + fp32 accumulation-order noise; fp32 heads are tighter
    tol = scale * (1e-3 if t["heads"] else 2 ** -7)
    assert torch.isfinite(out).all()
    err = float((out - ref).abs().max())
    assert err <= tol, f"{case[0]}: max abs err {err} > {tol} (scale {scale})"


def _fuzz_cases(count, seed):
    import random
    rnd = random.Random(seed)
    cases = []
    for i in range(count):
        ksize = rnd.choice([1, 1, 3])
        heads = ksize == 1 and rnd.random() < 0.15
        cin = rnd.choice([64, 128, 256]) if not heads else 256
        cout = rnd.randint(1, 32) if heads else rnd.choice([32, 64, 128, 256] if ksize == 1 else [64, 128])
        if ksize == 3:
            cin = cout
        up = (not heads) and ksize == 1 and rnd.random() < 0.3
        h, w = rnd.randint(1, 20), rnd.randint(1, 20)
        if up or rnd.random() < 0.5:
            h, w = 2 * h, 2 * w
        n = rnd.randint(1, 7)
        residual = up or ((not heads) and rnd.random() < 0.4)
        prologue = ksize == 1 and not heads and rnd.random() < 0.4
        cin2 = rnd.choice([0, 0, 64, 128]) if (ksize == 1 and not heads and not prologue) else 0
        cases.append((f"fuzz{i}", n, h, w, cin, cout, ksize, rnd.random() < 0.5, prologue, residual, up, cin2, heads))
    return cases


def test_conv_fuzz_random_shapes(ops, dev):
    """Seeded random shapes (ragged tiles, odd sizes, 1-pixel images, every epilogue combination): a case either
    matches the fp32 reference or is REJECTED with an HgError -- never a wrong answer, a hang or a sticky error."""
    from hgb200 import HgError
    ran, rejected = 0, []
    for case in _fuzz_cases(72, seed=11):
        t = make_conv_case(case, dev)
        try:
            out = run_conv_case(t)
        except (HgError, ValueError) as e:
            rejected.append((case, str(e)[:80]))
            continue
        ref = conv_reference(t)
        scale = float(ref.abs().max())
        tol = scale * (1e-3 if t["heads"] else 2 ** -7)
        err = float((out - ref).abs().max())
        assert torch.isfinite(out).all() and err <= tol, f"{case}: max abs err {err} > {tol}"
        ran += 1
    assert ran >= 48, f"only {ran} cases ran; rejected: {rejected}"


def test_conv_rejects_bad_shapes(ops, dev):
    from hgb200 import HgError
    x = torch.zeros(1, 8, 8, 48, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(64, 48, dtype=torch.bfloat16, device=dev)
    b = torch.zeros(64, device=dev)
    with pytest.raises(HgError):
        ops.conv_nhwc(x, w, b, ksize=1, cout=64)          # cin not a multiple of 64


def test_stem_im2col_and_gemm(ops, dev):
    no_tf32()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 64, 96, generator=g)
    wt = torch.randn(64, 3, 7, 7, generator=g) / 12.0
    bias = torch.randn(64, generator=g) * 0.1
    for flip in (False, True):
        rows = ops.stem_im2col(x.to(dev), flip_w=flip)
        assert rows.shape == (2, 32, 48, 192)
        wmat = torch.zeros(64, 192)
        wmat[:, :147] = wt.permute(0, 2, 3, 1).reshape(64, 147)
        out = ops.conv_nhwc(rows, wmat.to(torch.bfloat16).to(dev), bias.to(dev), ksize=1, cout=64, relu=True)
        torch.cuda.synchronize()
        ops.check_err_word(dev)
        xin = x.flip(-1) if flip else x
        ref = F.relu(F.conv2d(r16(xin).to(dev), r16(wt).to(dev), bias.to(dev), stride=2, padding=3))
        err = float((out.float().permute(0, 3, 1, 2) - ref).abs().max())
        assert err <= float(ref.abs().max()) * 2 ** -7


@pytest.mark.parametrize("shape", [(2, 64, 64, 128, 128), (3, 32, 32, 128, 128), (1, 128, 128, 64, 64), (5, 16, 16, 128, 128),
                                   (2, 64, 48, 128, 128), (7, 8, 8, 64, 128), (40, 64, 64, 128, 128), (3, 10, 6, 128, 64)])
def test_conv3x3_halo(shape, ops, dev):
    """Second-generation 3x3: halo-padded input, every tap a shifted descriptor into ONE smem tile."""
    no_tf32()
    n, h, w, cin, cout = shape
    g = torch.Generator().manual_seed(21)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cin ** 0.5))
    bias = (torch.randn(cout, generator=g) * 0.5).to(dev)
    buf = ops.halo_padded_buffer(n, h, w, cin, dev)
    ops.halo_interior(buf, n, h, w, cin).copy_(x.to(dev))
    wmat = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).to(torch.bfloat16).to(dev)
    out = ops.conv3x3_halo(buf, wmat, bias, n=n, h=h, w=w, cin=cin, cout=cout, relu=True)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    ref = F.relu(F.conv2d(x.float().to(dev).permute(0, 3, 1, 2), r16(wt).to(dev), bias, padding=1))
    err = float((out.float().permute(0, 3, 1, 2) - ref).abs().max())
    assert err <= float(ref.abs().max()) * 2 ** -7, f"err {err} vs scale {float(ref.abs().max())}"


def _halo_case(n, h, w, dev, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, h, w, 128, generator=g).to(torch.bfloat16)
    buf = None
    from hgb200 import ops as _ops
    buf = _ops.halo_padded_buffer(n, h, w, 128, dev)
    _ops.halo_interior(buf, n, h, w, 128).copy_(x.to(dev))
    w2 = (torch.randn(128, 9 * 128, generator=g) / (3.0 * 128 ** 0.5)).to(torch.bfloat16).to(dev)
    b2 = (torch.randn(128, generator=g) * 0.5).to(dev)
    return g, buf, w2, b2


@pytest.mark.parametrize("shape", [(40, 64, 64), (64, 64, 48), (300, 16, 16), (41, 30, 22)])
def test_conv3x3_paired_cta_kernel_is_bit_identical(shape, ops, dev, monkeypatch):
    """cta_group::2 variant of the halo 3x3 (HG_CONV3X3_PAIR): same MMAs in the same order -> the same bits."""
    n, h, w = shape
    _, buf, w2, b2 = _halo_case(n, h, w, dev, seed=31)
    monkeypatch.delenv("HG_CONV3X3_PAIR", raising=False)
    want = ops.conv3x3_halo(buf, w2, b2, n=n, h=h, w=w, cin=128, cout=128, relu=True)
    monkeypatch.setenv("HG_CONV3X3_PAIR", "1")
    got = ops.conv3x3_halo(buf, w2, b2, n=n, h=h, w=w, cin=128, cout=128, relu=True)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    assert torch.equal(got, want)


@pytest.mark.parametrize("shape", [(40, 64, 64, True, False), (40, 64, 64, True, True), (64, 64, 48, True, True),
                                   (300, 16, 16, True, False), (41, 30, 22, False, False), (90, 32, 32, False, True)])
def test_conv3x3_k3_fused_matches_the_two_kernels(shape, ops, dev):
    """Bottleneck tail in one launch (3x3 -> bias/ReLU -> 1x1 128->256 + residual (+ upsample-add)) on CTA pairs:
    bit-identical to hg_conv3x3_halo_bf16 followed by hg_conv_nhwc_bf16 (same roundings, same accumulation order)."""
    n, h, w, use_res, use_up = shape
    g, buf, w2, b2 = _halo_case(n, h, w, dev, seed=41)
    w3 = (torch.randn(256, 128, generator=g) / 128 ** 0.5).to(torch.bfloat16).to(dev)
    b3 = (torch.randn(256, generator=g) * 0.5).to(dev)
    res = torch.randn(n, h, w, 256, generator=g).to(torch.bfloat16).to(dev) if use_res else None
    up = torch.randn(n, h // 2, w // 2, 256, generator=g).to(torch.bfloat16).to(dev) if use_up else None
    assert ops.conv3x3_k3_fusable(n, h, w)
    z2 = ops.conv3x3_halo(buf, w2, b2, n=n, h=h, w=w, cin=128, cout=128, relu=True)
    want = ops.conv_nhwc(z2, w3, b3, ksize=1, cout=256, residual=res, up_low=up)
    got = ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=n, h=h, w=w, residual=res, up_low=up)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    assert torch.equal(got, want)


def test_conv3x3_k3_fused_rejects_what_it_cannot_run(ops, dev):
    from hgb200 import HgError
    assert not ops.conv3x3_k3_fusable(2, 8, 8)          # too few tiles for the CTA pairs: the caller keeps the two kernels
    _, buf, w2, b2 = _halo_case(40, 63, 64, dev, seed=1)
    w3 = torch.zeros(256, 128, dtype=torch.bfloat16, device=dev)
    b3 = torch.zeros(256, device=dev)
    up = torch.zeros(40, 31, 32, 256, dtype=torch.bfloat16, device=dev)
    with pytest.raises(HgError):
        ops.conv3x3_k3_fused(buf, w2, b2, w3, b3, n=40, h=63, w=64, up_low=up)        # odd height with an upsample operand


@pytest.mark.parametrize("shape", [(2, 64, 64, 256, 128), (3, 32, 32, 256, 128), (1, 128, 128, 64, 64), (4, 16, 16, 256, 128)])
def test_conv1x1_writes_halo_padded_output(shape, ops, dev):
    no_tf32()
    n, h, w, cin, cout = shape
    g = torch.Generator().manual_seed(22)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(dev)
    wt = (torch.randn(cout, cin, generator=g) / cin ** 0.5).to(torch.bfloat16).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.5).to(dev)
    scale = (0.5 + torch.rand(cin, generator=g)).to(dev)
    shift = (0.3 * torch.randn(cin, generator=g)).to(dev)
    dense = ops.conv_nhwc(x, wt, bias, ksize=1, cout=cout, relu=True, in_scale=scale, in_shift=shift)
    buf = ops.halo_padded_buffer(n, h, w, cout, dev)
    ops.conv_nhwc(x, wt, bias, ksize=1, cout=cout, relu=True, in_scale=scale, in_shift=shift, out_halo=buf)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    assert torch.equal(ops.halo_interior(buf, n, h, w, cout), dense)
    full = buf[(w + 1) * cout:].view(n, h + 1, w + 1, cout)
    assert float(full[:, h].abs().max()) == 0 and float(full[:, :, w].abs().max()) == 0      # pads untouched
    assert float(buf[:(w + 1) * cout].abs().max()) == 0


def test_halo_chain_fuzz_random_shapes(ops, dev):
    """1x1 (+bn prologue) -> halo-padded buffer -> 3x3 on seeded random shapes (odd widths, 1-row images, widths up
    to the kernel's limit): bit-identical to the dense 1x1 followed by the generic 3x3 path's operands, and within one
    bf16 rounding of the fp32 reference."""
    import random
    from hgb200 import HgError
    no_tf32()
    rnd = random.Random(5)
    ran = 0
    for i in range(40):
        n, h = rnd.randint(1, 6), rnd.randint(1, 40)
        w = rnd.choice([1, 2, 3, 5, 8, 17, 31, 48, 64, 100, 253]) if i % 4 == 0 else rnd.randint(1, 72)
        cin, c = rnd.choice([64, 128, 256]), rnd.choice([64, 128])
        if w > 100:
            n, h = 1, rnd.randint(1, 6)
        g = torch.Generator().manual_seed(100 + i)
        x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(dev)
        w1 = (torch.randn(c, cin, generator=g) / cin ** 0.5).to(torch.bfloat16).to(dev)
        b1 = (torch.randn(c, generator=g) * 0.5).to(dev)
        w3 = torch.randn(c, c, 3, 3, generator=g) / (3.0 * c ** 0.5)
        b3 = (torch.randn(c, generator=g) * 0.5).to(dev)
        scale = (0.5 + torch.rand(cin, generator=g)).to(dev)
        shift = (0.3 * torch.randn(cin, generator=g)).to(dev)
        try:
            dense = ops.conv_nhwc(x, w1, b1, ksize=1, cout=c, relu=True, in_scale=scale, in_shift=shift)
            buf = ops.halo_padded_buffer(n, h, w, c, dev)
            ops.conv_nhwc(x, w1, b1, ksize=1, cout=c, relu=True, in_scale=scale, in_shift=shift, out_halo=buf)
            wmat = w3.permute(0, 2, 3, 1).reshape(c, 9 * c).to(torch.bfloat16).to(dev)
            out = ops.conv3x3_halo(buf, wmat, b3, n=n, h=h, w=w, cin=c, cout=c, relu=bool(i & 1))
        except (HgError, ValueError):
            continue
        torch.cuda.synchronize()
        ops.check_err_word(dev)
        assert torch.equal(ops.halo_interior(buf, n, h, w, c), dense), (n, h, w, cin, c)
        ref = F.conv2d(dense.float().permute(0, 3, 1, 2), r16(w3).to(dev), b3, padding=1)
        if i & 1:
            ref = F.relu(ref)
        err = float((out.float().permute(0, 3, 1, 2) - ref).abs().max())
        assert err <= float(ref.abs().max()) * 2 ** -7, f"{(n, h, w, cin, c)}: err {err} vs scale {float(ref.abs().max())}"
        ran += 1
    assert ran >= 30, ran


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 256, 256), (2, 256, 192), (3, 128, 512)])
def test_stem_window_conv(shape, ops, dev):
    """7x7/s2 stem through the overlapping-window TMA path (no im2col matrix)."""
    from hgb200.fold import stem_window_weight
    no_tf32()
    n, h, w = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, 3, h, w, generator=g)
    wt = torch.randn(64, 3, 7, 7, generator=g) / 12.0
    bias = torch.randn(64, generator=g) * 0.1
    wwin = stem_window_weight(wt).to(dev)
    for flip in (False, True):
        packed = ops.stem_packed_buffer(n, h, w, dev)
        ops.stem_pack(x.to(dev), packed, flip_w=flip)
        xin = x.flip(-1) if flip else x
        assert torch.equal(packed[:, :, 4:-4, :3], xin.permute(0, 2, 3, 1).to(torch.bfloat16).to(dev))
        assert float(packed[:, :, :4].abs().max()) == 0 and float(packed[:, :, -4:].abs().max()) == 0
        out = ops.stem_conv(packed, wwin, bias.to(dev))
        torch.cuda.synchronize()
        ops.check_err_word(dev)
        ref = F.relu(F.conv2d(r16(xin).to(dev), r16(wt).to(dev), bias.to(dev), stride=2, padding=3))
        err = float((out.float().permute(0, 3, 1, 2) - ref).abs().max())
        assert err <= float(ref.abs().max()) * 2 ** -7, f"err {err}"
    # both orientations of the flip test from one read: [images | mirrors]
    a, b, both = (ops.stem_packed_buffer(k, h, w, dev) for k in (n, n, 2 * n))
    ops.stem_pack(x.to(dev), a, flip_w=False)
    ops.stem_pack(x.to(dev), b, flip_w=True)
    ops.stem_pack(x.to(dev), both, flip_w="both")
    assert torch.equal(both[:n], a) and torch.equal(both[n:], b)


@pytest.mark.parametrize("shape", [(2, 64, 64, 256), (3, 32, 32, 256), (8, 16, 16, 256), (2, 64, 64, 128), (5, 32, 32, 256), (1, 8, 8, 256)])
def test_conv1x1_prologue_also_writes_the_pooled_input(shape, ops, dev):
    """hg_conv_desc.pool_in: the 2x2 max-pool of the RAW input (what Hourglass pools, src/models/modules.py:82) as a second
    output of the prologue warps -- bit-identical to the pool kernel, and the conv's own halo-padded result unchanged."""
    from hgb200 import HgError
    n, h, w, cin = shape
    cout = 128
    g = torch.Generator().manual_seed(31)
    x = torch.randn(n, h, w, cin, generator=g).to(torch.bfloat16).to(dev)
    wt = (torch.randn(cout, cin, generator=g) / cin ** 0.5).to(torch.bfloat16).to(dev)
    bias = (torch.randn(cout, generator=g) * 0.5).to(dev)
    scale = (0.5 + torch.rand(cin, generator=g)).to(dev)
    shift = (0.3 * torch.randn(cin, generator=g)).to(dev)
    assert ops.conv_pool_in_fusable(n, h, w, cout) == ((n * h * w) % 128 == 0)
    if not ops.conv_pool_in_fusable(n, h, w, cout):
        with pytest.raises(HgError):
            ops.conv_nhwc(x, wt, bias, ksize=1, cout=cout, relu=True, in_scale=scale, in_shift=shift,
                          pool_in=torch.empty(n, h // 2, w // 2, cin, dtype=torch.bfloat16, device=dev))
        return
    plain = ops.halo_padded_buffer(n, h, w, cout, dev)
    ops.conv_nhwc(x, wt, bias, ksize=1, cout=cout, relu=True, in_scale=scale, in_shift=shift, out_halo=plain)
    both = ops.halo_padded_buffer(n, h, w, cout, dev)
    pooled = torch.full((n, h // 2, w // 2, cin), -7.0, dtype=torch.bfloat16, device=dev)
    ops.conv_nhwc(x, wt, bias, ksize=1, cout=cout, relu=True, in_scale=scale, in_shift=shift, out_halo=both, pool_in=pooled)
    torch.cuda.synchronize()
    ops.check_err_word(dev)
    assert torch.equal(both, plain)
    assert torch.equal(pooled, ops.maxpool2x2(x))
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 2, stride=2).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(pooled, ref)


# ------------------------------------------------------------------------------------------ bandwidth kernels
@pytest.mark.parametrize("shape", [(2, 64, 64, 256), (3, 8, 8, 128), (1, 4, 6, 64), (5, 2, 2, 256)])
def test_maxpool_bit_exact(shape, ops, dev):
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
    out = ops.maxpool2x2(x)
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 2, stride=2).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("shape", [(2, 64, 64, 256), (3, 8, 8, 128), (1, 4, 6, 64)])
def test_upsample_add_bit_exact(shape, ops, dev):
    g = torch.Generator().manual_seed(2)
    n, h, w, c = shape
    a = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16).to(dev)
    low = torch.randn(n, h // 2, w // 2, c, generator=g).to(torch.bfloat16).to(dev)
    out = ops.upsample2x_add(a, low)
    up = F.interpolate(low.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    ref = (a.float() + up).to(torch.bfloat16)
    assert torch.equal(out, ref)


def test_bn_relu(ops, dev):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 7, 5, 128, generator=g).to(torch.bfloat16).to(dev)
    s = (0.5 + torch.rand(128, generator=g)).to(dev)
    t = torch.randn(128, generator=g).to(dev)
    out = ops.bn_relu(x, s, t)
    ref = F.relu(torch.addcmul(t, x.float(), s)).to(torch.bfloat16)      # fma, one rounding
    assert (out.float() - ref.float()).abs().max() <= 2 ** -7 * ref.float().abs().max()
    assert (out == ref).float().mean() > 0.99


def test_layout_roundtrip(ops, dev):
    x = torch.randn(2, 24, 6, 10, generator=torch.Generator().manual_seed(5)).to(dev)
    nhwc = ops.nchw_to_nhwc_bf16(x)
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).to(torch.bfloat16))
    back = ops.nhwc_bf16_to_nchw(nhwc)
    assert torch.equal(back, x.to(torch.bfloat16).float())


# ------------------------------------------------------------------------------------------ decode (bit-exact)
def test_decode_argmax_golden_bit_exact(ops, dev):
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    for i, hm in enumerate(GI.heatmap_cases()):
        preds, maxval, idx = ops.decode_argmax(torch.from_numpy(hm).to(dev))
        np.testing.assert_array_equal(preds.cpu().numpy(), z[f"preds{i}"])          # reference output
        np.testing.assert_array_equal(preds.cpu().numpy(), D.get_preds(hm))         # oracle
        flat = hm.reshape(hm.shape[0], hm.shape[1], -1)
        np.testing.assert_array_equal(idx.cpu().numpy(), flat.argmax(2).astype(np.int32))
        np.testing.assert_array_equal(maxval.cpu().numpy(), flat.max(2))


def test_decode_argmax_random_with_ties(ops, dev):
    rng = np.random.RandomState(0)
    for (b, j, h, w) in [(128, 16, 64, 64), (7, 17, 64, 48), (3, 5, 7, 9), (2, 3, 1, 1), (64, 21, 16, 16)]:
        # few distinct values -> many ties; first flat index must win
        hm = rng.randint(-3, 4, size=(b, j, h, w)).astype(np.float32)
        preds, _, idx = ops.decode_argmax(torch.from_numpy(hm).to(dev))
        np.testing.assert_array_equal(idx.cpu().numpy(), hm.reshape(b, j, -1).argmax(2).astype(np.int32))
        np.testing.assert_array_equal(preds.cpu().numpy(), D.get_preds(hm))


def test_decode_final_preds_golden(ops, dev):
    z = np.load(os.path.join(GOLDEN, "decode.npz"))
    cases = GI.heatmap_cases()
    centers, scales = GI.decode_args(cases)
    for i, hm in enumerate(cases):
        B, J, H, W = hm.shape
        out = ops.decode_final_preds(torch.from_numpy(hm).to(dev), centers[i], scales[i], (W, H)).cpu().numpy()
        # quarter-pixel offsets are exact; the float64 affine is solved by Cramer instead of cv2's LU
        np.testing.assert_allclose(out, z[f"final{i}"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(out, D.get_final_preds_batch(hm, centers[i], scales[i], (W, H)), rtol=0, atol=1e-9)


def test_flip_average(ops, dev):
    rng = np.random.RandomState(1)
    for J, pairs, (H, W) in [(16, D.MPII_FLIP_PAIRS, (64, 64)), (17, D.COCO_FLIP_PAIRS, (64, 48))]:
        hm = rng.randn(3, J, H, W).astype(np.float32)
        hf = rng.randn(3, J, H, W).astype(np.float32)
        perm = torch.from_numpy(D.flip_perm(J, pairs).astype(np.int32)).to(dev)
        out = ops.flip_average(torch.from_numpy(hm).to(dev), torch.from_numpy(hf).to(dev), perm)
        np.testing.assert_array_equal(out.cpu().numpy(), D.flip_average(hm, hf, pairs))


def test_pck_dists(ops, dev):
    for pred, tgt in GI.accuracy_cases():
        d = ops.pck_dists(torch.from_numpy(pred).to(dev), torch.from_numpy(tgt).to(dev)).cpu().numpy()
        B, J = pred.shape[:2]
        p, g = D.get_preds(pred), D.get_preds(tgt)
        norm = np.float32(pred.shape[3]) / np.float32(10)
        for b in range(B):
            for j in range(J):
                if g[b, j, 0] > 1 and g[b, j, 1] > 1:
                    e = p[b, j] - g[b, j]
                    ref = np.float32(np.sqrt(np.float32(e[0] * e[0] + e[1] * e[1]))) / norm
                    assert abs(d[j, b] - ref) <= 1e-6 * max(1.0, abs(ref))
                else:
                    assert d[j, b] == -1


# ------------------------------------------------------------------------------------------ targets + loss
def test_gaussian_target_bit_exact(ops, dev):
    z = np.load(os.path.join(GOLDEN, "loss.npz"))
    for i, c in enumerate(GI.loss_cases()):
        joints = torch.from_numpy(c["joints"]).to(dev)
        vis = torch.from_numpy(c["vis"]).to(dev)
        mu, wt = ops.joint_centers(joints, vis, c["hsz"], c["isz"], sigma=1)
        tgt = ops.gaussian_target(mu, wt, c["hsz"], sigma=1)
        np.testing.assert_array_equal(wt.cpu().numpy()[..., None], z[f"tw{i}"])
        np.testing.assert_array_equal(tgt.cpu().numpy(), z[f"target{i}"])


@pytest.mark.parametrize("on_the_fly", [False, True])
def test_jmse_loss_and_grad(on_the_fly, ops, dev):
    z = np.load(os.path.join(GOLDEN, "loss.npz"))
    stride = int(z["grad_stride"])
    for i, c in enumerate(GI.loss_cases()):
        tg, tw = z[f"target{i}"], z[f"tw{i}"]
        preds = [torch.from_numpy(tg + n).to(dev) for n in c["noise"]]
        twd = torch.from_numpy(tw).to(dev)
        if on_the_fly:
            mu, wt = ops.joint_centers(torch.from_numpy(c["joints"]).to(dev), torch.from_numpy(c["vis"]).to(dev),
                                       c["hsz"], c["isz"], sigma=1)
            loss, grads = ops.jmse_loss(preds, None, wt, want_grad=True, mu=mu, sigma=1)
        else:
            loss, grads = ops.jmse_loss(preds, torch.from_numpy(tg).to(dev), twd, want_grad=True)
        ref = float(z[f"loss{i}"])
        assert abs(float(loss.item()) - ref) <= 2e-5 * abs(ref)          # fp32 sums, different order
        oloss, ograds = L.joints_mse([p.cpu().numpy() for p in preds], tg, tw, True)
        assert abs(float(loss.item()) - oloss) <= 2e-5 * abs(oloss)
        for s, g in enumerate(grads):
            np.testing.assert_allclose(g.cpu().numpy().reshape(-1)[::stride], z[f"grad{i}_{s}"], rtol=2e-5, atol=1e-10)
    # use_target_weight=False
    c = GI.loss_cases()[0]
    tg = z["target0"]
    preds = [torch.from_numpy(tg + n).to(dev) for n in c["noise"]]
    loss, _ = ops.jmse_loss(preds, torch.from_numpy(tg).to(dev), None, want_grad=False)
    assert abs(float(loss.item()) - float(z["loss_nw0"])) <= 2e-5 * float(z["loss_nw0"])
